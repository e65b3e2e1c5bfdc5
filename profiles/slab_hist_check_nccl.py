"""2-rank check of the slab path's all-reduced histograms against the single-GPU entry point.
torchrun --nproc-per-node 2 profiles/slab_hist_check_nccl.py [size]"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-feature-extraction_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench, ife_b200
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("gloo")
ctx = ife_b200.Context(rank)
uid = [ctx.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(uid[0], world, rank)
size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
gd = (size, size, size)
full = bench.synth_volume_torch(torch, dev, 4, gd, z_range=(0, size))
fmask = torch.ones(gd, dtype=torch.uint8, device=dev)
z0, z1 = ife_b200.slab_range(size, world, rank)
img, mask = full[z0:z1], fmask[z0:z1]
S = bench.SIGMAS
edges = bench.equalized_edges_from_scan(torch, ctx, full[:64].contiguous(), fmask[:64].contiguous(), 40, dims=(size, size, 64))
box = [edges if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
edges = box[0]
rows, nb = 32, 41
ref = torch.zeros((1, rows, nb), dtype=torch.int32, device=dev)
ctx.emphysema_histograms_dev(full.data_ptr(), fmask.data_ptr(), ref.data_ptr(), gd, S, edges, None)
out = torch.empty((1, 8, z1 - z0, size, size), dtype=torch.float32, device=dev)
for mode in ("per_scale", "all"):
    cnt = torch.zeros((rows, nb), dtype=torch.int32, device=dev)
    if mode == "per_scale":
        for si, s in enumerate(S):
            ctx.slab_emphysema_features_dev(img.data_ptr(), mask.data_ptr(), out.data_ptr(), gd, [s], edges=edges[8*si:8*si+8],
                                            counts_ptr=cnt[8*si:].data_ptr(), halo_factor=1000.0)
    else:
        ctx.slab_emphysema_features_dev(img.data_ptr(), mask.data_ptr(), None, gd, S, edges=edges, counts_ptr=cnt.data_ptr(), halo_factor=1000.0)
    ctx.synchronize(); torch.cuda.synchronize()
    d = (cnt.to(torch.int64) - ref[0].to(torch.int64)).abs()
    print(rank, mode, "hist diff", int(d.sum()), "per row", d.sum(1).tolist(), "row sums", cnt.to(torch.int64).sum(1)[:4].tolist(), ref[0].to(torch.int64).sum(1)[:2].tolist(), flush=True)
ctx.comm_destroy(); ctx.close(); dist.destroy_process_group()
