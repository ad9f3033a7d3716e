// Shared declarations for the marf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/marf_b200.h"

namespace marf {

constexpr int kMaxBands = 16;          // posenc bands supported (arch.posenc.L_2D <= 16)

// ---------------------------------------------------------------------------------------------
// Geometry constants of one handle (warp.py:9-21) + encoding schedule, passed by value to kernels.
struct Geo {
  int H, W;                // canvas
  int h, w;                // patch grid actually traversed (patch_H x patch_W when cropped, else H x W)
  int y0, x0;              // crop origin (0 when not cropped)
  int rows, row_offset;    // local row shard of each patch
  int patch_offset;        // global index of local patch 0
  float norm_h, norm_w;    // H/max(H,W), W/max(H,W)
  int L;                   // posenc bands
  int d_in;                // 2 + 4L (or 2)
  float band_w[kMaxBands]; // c2f weights w_k (1 when c2f is off)
  float band_f[kMaxBands]; // f32(2^k * pi)
};

// normalized pixel coordinate of (row r, col c) of the traversed grid — same f32 op order as
// warp.py:38-49: ((i + 0.5) / n * 2 - 1) * norm
__device__ __forceinline__ void grid_xy(const Geo& g, int r, int c, float& x, float& y) {
  float fy = (float)(r + g.y0) + 0.5f;
  float fx = (float)(c + g.x0) + 0.5f;
  y = (__fdiv_rn(fy, (float)g.H) * 2.0f - 1.0f) * g.norm_h;
  x = (__fdiv_rn(fx, (float)g.W) * 2.0f - 1.0f) * g.norm_w;
}

// homography apply, warp.py:74-78: q = [x,y,1] H^T ; (u,v) = q_xy / (q_z + 1e-8)
__device__ __forceinline__ void apply_h(const float* __restrict__ Hm, float x, float y, float& u, float& v, float& qz) {
  float q0 = Hm[0] * x + Hm[1] * y + Hm[2];
  float q1 = Hm[3] * x + Hm[4] * y + Hm[5];
  float q2 = Hm[6] * x + Hm[7] * y + Hm[8];
  qz = q2 + 1e-8f;
  u = __fdiv_rn(q0, qz);
  v = __fdiv_rn(q1, qz);
}

struct PxRange {
  long long first;   // first local pixel-sample of the chunk (index into batch*rows*w)
  int count;         // valid pixel-samples in the chunk
  int padded;        // rows allocated (multiple of 128); rows >= count are written as zeros
};

__device__ __forceinline__ void decode_px(const Geo& g, long long i, int& b, int& r, int& c) {
  const long long per = (long long)g.rows * g.w;
  int rem;
  if (i < 0x7fffffffLL) {                     // (64-bit division is ~10x the cost of the 32-bit one; the branch is uniform)
    const unsigned ui = (unsigned)i, up = (unsigned)per;
    b = (int)(ui / up);
    rem = (int)(ui - (unsigned)b * up);
  } else {
    b = (int)(i / per);
    rem = (int)(i - (long long)b * per);
  }
  const int q = rem / g.w;
  r = q + g.row_offset;
  c = rem - q * g.w;
}

// Programmatic dependent launch: block until every kernel this launch depends on has completed and flushed.
// A no-op for launches without the programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Every kernel of the step is launched with programmatic stream serialization (PDL): the next kernel's launch and
// prologue overlap the previous kernel's tail; each kernel calls griddepcontrol.wait before it touches dependent data.
// MARF_NO_PDL=1 falls back to plain stream-ordered launches.
inline bool use_pdl() {
  static const bool on = getenv("MARF_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same with a thread-block cluster of `cluster` CTAs along x
template <typename... KArgs, typename... Args>
inline void launch_k_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int pad4(int a) { return (a + 3) / 4 * 4; }

// loss coefficients resolved on device from the (possibly all-reduced) sums: no host sync.
struct LossCoef {
  double inv_n_rgb;    // 1 / N_RGB
  double s_over_n2;    // 3 * S_RGB / N_RGB^2        (d rgb_loss / d m, constant part)
  double inv_n_mask;   // 1 / N_MASK
  double inv_n_edge;   // 1 / N_EDGE  (0 when no edge term)
  double se_over_n2;   // 3 * S_EDGE / N_EDGE^2
};

}  // namespace marf
