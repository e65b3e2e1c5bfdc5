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
        # north_star terms: voxels whose eigenvalues differ by more than 1e-4 * max|lambda|
        lam, lam_ref = out[:, 2:5].astype(np.float64), ref[:, 2:5].astype(np.float64)
        scale = np.abs(lam_ref).max(1)
        rel = np.where(scale > 0, np.abs(lam - lam_ref).max(1) / np.where(scale > 0, scale, 1), 0.0)
        bad = torch.tensor([float((rel > 1e-4).sum()), float(rel.size)], dtype=torch.float64)
        dist.all_reduce(bad, op=dist.ReduceOp.SUM)
        res[name] = {"max_mismatching_values_per_rank": stats[0].item(), "max_abs_diff": stats[1].item(),
                     "hist_abs_diff_vs_whole": stats[2].item(), "values_per_rank": int(out.size),
                     "frac_voxels_eig_err_gt_1e-4": bad[0].item() / bad[1].item(),
                     "hist_inserts": int(whole_counts.sum())}
        if name == "full_halo":
            # two device-mode calls with histograms back to back, no synchronisation in between:
            # the second call's halo exchange (copy stream) must order itself behind the first
            # call's all-reduce (main stream) -- one communicator, never two streams at once
            dev = torch.device("cuda", local)
            d_img = torch.from_numpy(img[z0:z1].copy()).to(dev)
            d_mask = torch.from_numpy(mask[z0:z1].copy()).to(dev)
            d_out = [torch.empty((2, 8) + d_img.shape, dtype=torch.float32, device=dev) for _ in range(2)]
            d_cnt = [torch.zeros(counts.shape, dtype=torch.int32, device=dev) for _ in range(2)]
            torch.cuda.synchronize()
            for it in range(2):
                ctx.slab_emphysema_features_dev(d_img.data_ptr(), d_mask.data_ptr(), d_out[it].data_ptr(),
                                                (shape[2], shape[1], nz), sigmas, edges=edges,
                                                counts_ptr=d_cnt[it].data_ptr(), halo_factor=hf)
            ctx.synchronize()
            torch.cuda.synchronize()
            ok = all(np.array_equal(d_out[it].cpu().numpy(), out, equal_nan=True) and
                     np.array_equal(d_cnt[it].cpu().numpy().astype(np.int64), counts.astype(np.int64)) for it in range(2))
            okt = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64)
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            res["back_to_back_device_calls_equal_host_call"] = bool(okt.item() == 1.0)
    ctx.comm_destroy()
    ctx.close()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps({"world": world, **res}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
