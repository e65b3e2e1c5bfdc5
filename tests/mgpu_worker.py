"""torchrun worker: z-slab path over NCCL (ife_cuda_slab_emphysema_features) against the
single-GPU entry point.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "image-feature-extraction_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import ife_b200
import synth


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = ife_b200.Context(local)
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], world, rank)
    shape = (64 * world, 48, 56)
    nz = shape[0]
    sigmas = [0.6, 2.4]
    img = synth.ct_like(shape, seed=31, n_blobs=30)
    mask = synth.clamp01(synth.lung_mask(shape))
    z0, z1 = ife_b200.slab_range(nz, world, rank)
    res = {}
    for hf, name in ((0.0, "default_halo"), (1000.0, "full_halo")):
        # reference on every rank (cheap at this size) for the comparison
        whole = ctx.emphysema_features(img, mask, sigmas)
        edges = np.stack([synth.equalized_edges(whole[s, k][mask != 0], 20) for s in range(2) for k in range(8)])
        whole_counts = ctx.emphysema_histograms(img, mask, sigmas, edges)[0]
        out, counts = ctx.slab_emphysema_features(img[z0:z1], mask[z0:z1], (shape[2], shape[1], nz), sigmas,
                                                  edges=edges, halo_factor=hf)
        ref = whole[:, :, z0:z1]
        neq = ~((out == ref) | (np.isnan(out) & np.isnan(ref)))
        stats = torch.tensor([float(neq.sum()), float(np.abs(out.astype(np.float64) - ref)[neq].max() if neq.any() else 0.0),
                              float(np.abs(counts.astype(np.int64) - whole_counts).sum())], dtype=torch.float64)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        res[name] = {"max_mismatching_values_per_rank": stats[0].item(), "max_abs_diff": stats[1].item(),
                     "hist_abs_diff_vs_whole": stats[2].item(), "values_per_rank": int(out.size)}
    ctx.comm_destroy()
    ctx.close()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps({"world": world, **res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
