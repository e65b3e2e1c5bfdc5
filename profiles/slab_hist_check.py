import os, sys, numpy as np
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT,'image-feature-extraction_b200')); sys.path.insert(0, os.path.join(ROOT,'tests'))
import torch, ife_b200, synth
ctx=ife_b200.Context(0)
for shape in ((128,256,256),(128,1024,1024),(600,512,512)):
    nz,ny,nx=shape
    img=synth.ct_like(shape, seed=3, n_blobs=40)
    mask=np.ones(shape,np.uint8)
    sigmas=[0.6,4.8]
    whole=ctx.emphysema_features(img,mask,sigmas)
    edges=np.stack([synth.equalized_edges(whole[s,k].reshape(-1)[::7],40) for s in range(2) for k in range(8)])
    wc=ctx.emphysema_histograms(img,mask,sigmas,edges)[0].astype(np.int64)
    tot=np.zeros_like(wc)
    for r in range(2):
        z0,z1=ife_b200.slab_range(nz,2,r)
        out,cnt=ctx.slab_emphysema_features_local(img,mask,0,z0,z1-z0,(nx,ny,nz),sigmas,edges=edges,halo_factor=1000.0)
        print(shape,'rank',r,'values differing', int((out!=whole[:,:,z0:z1]).sum()))
        tot+=cnt.astype(np.int64)
    d=np.abs(tot-wc)
    print(shape,'hist diff total',d.sum(),'per row',d.sum(1).tolist(), 'row sums', tot.sum(1)[:3].tolist(), wc.sum(1)[:3].tolist())
