// Internal: the per-GPU context behind the opaque ife_cuda_ctx of include/ife_cuda.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

struct ife_cuda_ctx;

namespace ife {

// Grow-only device allocation owned by a context.
struct DeviceBuffer {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(ife_cuda_ctx* ctx, size_t want);
  void release();
};

struct Workspace {
  DeviceBuffer a0, a1, b0, b1;  // float volumes between the z, x and y passes (two fields)
  DeviceBuffer blur;            // smoothed volume fed to the fused feature kernel
  DeviceBuffer ckpt;            // recursion checkpoints (double)
  DeviceBuffer in_img, in_mask; // staging for IFE_MEM_HOST inputs
  DeviceBuffer out[2];          // staging for IFE_MEM_HOST outputs (double-buffered per scale)
  DeviceBuffer edges, rois, counts;
  DeviceBuffer slab_img, slab_mask;  // slab + halo planes (multi-GPU)
  void release_all() {
    DeviceBuffer* all[] = {&a0, &a1, &b0, &b1, &blur, &ckpt, &in_img, &in_mask, &out[0], &out[1],
                           &edges, &rois, &counts, &slab_img, &slab_mask};
    for (DeviceBuffer* b : all) b->release();
  }
};

int fail(ife_cuda_ctx* ctx, int code, const char* fmt, ...);

}  // namespace ife

struct ife_cuda_ctx {
  int device = 0;
  int sm_count = 148;
  int arith = 1;  // IFE_ARITH_FMA
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t user_stream = nullptr;
  bool use_user_stream = false;
  cudaEvent_t events[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  std::string error;
  ife::Workspace ws;
  // NCCL (resolved lazily with dlopen; see slab.cu)
  void* nccl_comm = nullptr;
  int n_ranks = 1;
  int rank = 0;

  cudaStream_t stream() const { return use_user_stream ? user_stream : own_stream; }
};
