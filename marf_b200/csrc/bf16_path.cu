// precision=bf16: the tensor-core path.  bf16 activations/weights, fp32 accumulation in TMEM (tcgen05.mma),
// fp32 master weights and gradients.  A step of the 256-wide networks is 10 launches: prologue (weight pack + homographies +
// mask class table + zeroing), encoding, k_tc_chain<fwd> (all layers of both MLPs on CTA pairs), loss statistics, edge
// branch, loss gradients, k_tc_chain<dx>, the dX0 / warp-gradient GEMM, k_tc_dw (every dW / db), tail (mask layer-0
// gradients + sl(3) adjoint + gradient hand-over).  512-wide networks and MARF_NO_FUSE run the per-layer k_tc_gemm launches
// with SIMT 3-/1-wide output layers.  No fallback to the fp32 path.
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "engine.cuh"
#include "fp32_kernels.cuh"
#include "tc_kernels.cuh"
#include "tc_chain.cuh"
#include "tc_bwd.cuh"

namespace marf {

using bf16 = __nv_bfloat16;

#define BF_TRY(h, expr)                                                                         \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) return fail(h, MARF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)
#define BF_LAUNCH(h)                                                                            \
  do {                                                                                          \
    (h)->launches++;                                                                            \
    cudaError_t e__ = cudaGetLastError();                                                       \
    if (e__ == cudaSuccess && getenv("MARF_DEBUG_SYNC")) {                                      \
      e__ = cudaDeviceSynchronize();                                                            \
      if (e__ != cudaSuccess) fprintf(stderr, "MARF_DEBUG_SYNC: fault after launch at %s:%d\n", __FILE__, __LINE__); \
    }                                                                                           \
    if (e__ != cudaSuccess) return fail(h, MARF_ERR_CUDA, std::string("bf16 kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------ SIMT helpers
template <int LT>   // LT > 0: compile-time band count (registers); LT == 0: runtime g.L (local array)
__global__ void k_encode_bf16(Geo g, PxRange rg, const float* __restrict__ Hm, bf16* __restrict__ X0, int ld) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rg.padded) return;
  uint4* o = reinterpret_cast<uint4*>(X0 + (size_t)t * ld);
  float f[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) f[j] = 0.f;
  if (t < rg.count) {
    int b, r, c;
    decode_px(g, rg.first + t, b, r, c);
    float x, y, u, v, qz;
    grid_xy(g, r, c, x, y);
    apply_h(Hm + 9 * (b + g.patch_offset), x, y, u, v, qz);
    f[0] = u; f[1] = v;
    const int L = LT > 0 ? LT : g.L;
#pragma unroll
    for (int k = 0; k < (LT > 0 ? LT : kMaxBands); ++k) {
      if (k < L && 2 + 3 * L + k < 64) {
        float su, cu, sv, cv;
        sincosf(u * g.band_f[k], &su, &cu);
        sincosf(v * g.band_f[k], &sv, &cv);
        float wk = g.band_w[k];
        f[2 + k] = su * wk; f[2 + L + k] = cu * wk; f[2 + 2 * L + k] = sv * wk; f[2 + 3 * L + k] = cv * wk;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = make_uint4(tc::pack_bf16(f[8 * j], f[8 * j + 1]), tc::pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                      tc::pack_bf16(f[8 * j + 4], f[8 * j + 5]), tc::pack_bf16(f[8 * j + 6], f[8 * j + 7]));
  for (int j = 8; j < ld / 8; ++j) o[j] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- mask head layer 0 through the colour-class table (model/planar.py:342-349).  trunc(rgb) of a [0,1] image is 0 or 1,
// so the 384-wide colour embedding takes one of 8 values per pixel: layer 0 = T[class] + W0[:,384:] . PosEmbedding(xy),
// with T[c] = b0 + W0[:, :384] . colE(c) rebuilt every step (8 x 256 entries) — algebraically identical to the dense layer.
// Feature row (64 bf16): [PosEmbedding(xy) (2+4F) | onehot(class) (8) | onehot(class) (8) | 0...].  The two one-hot groups
// multiply the hi and lo bf16 halves of the class table T that k_mask_table writes into the packed layer-0 weights, so
// the colour part needs no epilogue work, and in the dW GEMM the first one-hot group yields the per-class column sums.
__global__ void k_mask_uv_cls(Geo g, PxRange rg, const float* __restrict__ rgb, int n_freqs, bf16* __restrict__ UV,
                              double* __restrict__ bad) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rg.padded) return;
  uint4* o = reinterpret_cast<uint4*>(UV + (size_t)t * 64);
  float f[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) f[j] = 0.f;
  unsigned char cl = 0;
  if (t < rg.count) {
    int b, r, c;
    long long i = rg.first + t;
    decode_px(g, i, b, r, c);
    long long per = (long long)g.rows * g.w;
    long long rem = i - (long long)b * per;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      long long idx = (long long)rgb[((long long)b * 3 + ch) * per + rem];
      if (idx < 0 || idx > 1) atomicAdd(bad, 1.0);      // counted, reported through loss_sums[MARF_BAD_INDEX] (no host sync)
      cl |= (unsigned char)((idx & 1) << ch);
    }
    float x, y;
    grid_xy(g, r, c, x, y);
    f[0] = x; f[1] = y;
#pragma unroll
    for (int fi = 0; fi < 15; ++fi) {
      if (fi < n_freqs) {
        const float fr = (float)(1 << fi);
        f[2 + 4 * fi + 0] = sinf(fr * x); f[2 + 4 * fi + 1] = sinf(fr * y);
        f[2 + 4 * fi + 2] = cosf(fr * x); f[2 + 4 * fi + 3] = cosf(fr * y);
      }
    }
  }
  if (t < rg.count) {
    const int k_uv = 2 + 4 * n_freqs;
#pragma unroll
    for (int j = 0; j < 64; ++j)
      if (j == k_uv + cl || j == k_uv + 8 + cl) f[j] = 1.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = make_uint4(tc::pack_bf16(f[8 * j], f[8 * j + 1]), tc::pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                      tc::pack_bf16(f[8 * j + 4], f[8 * j + 5]), tc::pack_bf16(f[8 * j + 6], f[8 * j + 7]));
}

// T[c][j] = b0[j] + sum_ch sum_e W0[j][ch*E + e] * embed[bit_ch(c)][e], one warp per (class, output) pair, written as
// bf16 hi / lo halves into columns k_uv + c and k_uv + 8 + c of the packed forward weights Wk [256, 64] of layer 0
__device__ __forceinline__ void mask_table_block(int vblock, const float* __restrict__ W0, const float* __restrict__ b0,
                                                 const float* __restrict__ embed, int k_in, int k_out, int edim, int k_uv,
                                                 bf16* __restrict__ Wk, int ldk) {
  const int gw = (vblock * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= 8 * k_out) return;
  const int c = gw / k_out, j = gw - c * k_out;
  const float* w = W0 + (size_t)j * k_in;
  float acc = 0.f;
  for (int i = lane; i < 3 * edim; i += 32) acc = fmaf(w[i], embed[(size_t)((c >> (i / edim)) & 1) * edim + (i % edim)], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    const float t = acc + b0[j];
    const bf16 hi = __float2bfloat16(t);
    Wk[(size_t)j * ldk + k_uv + c] = hi;
    Wk[(size_t)j * ldk + k_uv + 8 + c] = __float2bfloat16(t - __bfloat162float(hi));
  }
}
__global__ void k_mask_table(const float* __restrict__ W0, const float* __restrict__ b0, const float* __restrict__ embed,
                             int k_in, int k_out, int edim, int k_uv, bf16* __restrict__ Wk, int ldk) {
  pdl_wait();
  mask_table_block(blockIdx.x, W0, b0, embed, k_in, k_out, edim, k_uv, Wk, ldk);
}

// Layer-0 gradient of the mask head from the dW tile X = dY0^T [uv | onehot | onehot] ([k_out, 64] fp32) — computed by the
// tail launch (k_unpack_table modes 1 / 2):
//   uv columns          -> dW0[:, 3E : 3E + k_uv]
//   S[c][j] = X[j][k_uv + c] (per-class column sums of dY0)
//   dW0[j][ch*E + e] = sum_c S[c][j] * embed[bit_ch(c)][e],   db0[j] = sum_c S[c][j]

// W fp32 [rows, cols] -> bf16 [prow, pcol] (zero padded), optionally transposed: out[c][r]
__global__ void k_pack_bf16(const float* __restrict__ W, int rows, int cols, bf16* __restrict__ out, int prow, int pcol,
                            int transpose) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= prow * pcol) return;
  int pr = i / pcol, pc = i - pr * pcol;
  int r = transpose ? pc : pr, c = transpose ? pr : pc;
  out[i] = __float2bfloat16((r < rows && c < cols) ? W[(size_t)r * cols + c] : 0.f);
}

// thin output layer (k_out <= 4): logits[row, o] = b[o] + sum_k X[row,k] W[o,k]      (fp32 out [n, 4])
// lane owns k = lane*8..+7 (+256 per extra chunk) for every row, so its weights live in registers.
template <int OUT, int KCH>
__global__ void k_thin_fwd(int n, const bf16* __restrict__ X, int ld, const float* __restrict__ W,
                           const float* __restrict__ bias, float* __restrict__ out) {
  pdl_wait();
  const int width = KCH * 256;
  const int lane = threadIdx.x & 31;
  float w[KCH][OUT][8];
#pragma unroll
  for (int c = 0; c < KCH; ++c)
#pragma unroll
    for (int o = 0; o < OUT; ++o)
#pragma unroll
      for (int e = 0; e < 8; ++e) w[c][o][e] = W[o * width + c * 256 + lane * 8 + e];
  float bo[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) bo[o] = bias[o];
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  constexpr int R = 8;                               // rows in flight per warp
  for (int row0 = warp * R; row0 < n; row0 += nwarps * R) {
    uint4 raw[R][KCH];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int c = 0; c < KCH; ++c)
        raw[i][c] = row0 + i < n ? *reinterpret_cast<const uint4*>(X + (size_t)(row0 + i) * ld + c * 256 + lane * 8)
                                 : make_uint4(0u, 0u, 0u, 0u);
    float acc[R][OUT];
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
      for (int o = 0; o < OUT; ++o) acc[i][o] = 0.f;
#pragma unroll
      for (int c = 0; c < KCH; ++c) {
        const uint32_t w4[4] = {raw[i][c].x, raw[i][c].y, raw[i][c].z, raw[i][c].w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xv = __uint_as_float(((w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu) << 16);
#pragma unroll
          for (int o = 0; o < OUT; ++o) acc[i][o] = fmaf(xv, w[c][o][e], acc[i][o]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int o = 0; o < OUT; ++o) acc[i][o] = warp_sum(acc[i][o]);
    if (lane < R && row0 + lane < n) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (lane == i) { a0 = acc[i][0]; a1 = OUT > 1 ? acc[i][OUT > 1 ? 1 : 0] : 0.f; a2 = OUT > 2 ? acc[i][OUT > 2 ? 2 : 0] : 0.f; }
      *reinterpret_cast<float4*>(out + (size_t)(row0 + lane) * 4) =
          make_float4(a0 + bo[0], OUT > 1 ? a1 + bo[OUT > 1 ? 1 : 0] : 0.f, OUT > 2 ? a2 + bo[OUT > 2 ? 2 : 0] : 0.f, 0.f);
    }
  }
}

// dY_prev[row,k] = (sum_o dl[row,o] W[o,k]) * relu_mask[row,k]     (bf16 out); thread owns a fixed k-octet;
// the mask is the 1-bit-per-activation map the forward epilogue wrote
template <int OUT>
__global__ void k_thin_dx(int n, int width, const float* __restrict__ dl, const float* __restrict__ W,
                          const uint32_t* __restrict__ bits, int bits_ld, bf16* __restrict__ dY, int ldy) {
  pdl_wait();
  const int per_row = width / 8;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int total_threads = gridDim.x * blockDim.x;         // multiple of per_row (host guarantees)
  const int k = (tid % per_row) * 8;
  float w[OUT][8];
#pragma unroll
  for (int o = 0; o < OUT; ++o)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[o][e] = W[o * width + k + e];
  for (int row = tid / per_row; row < n; row += total_threads / per_row) {
    const float4 d4 = *reinterpret_cast<const float4*>(dl + (size_t)row * 4);
    const float d[3] = {d4.x, d4.y, d4.z};
    // layout: bit t <- column 2t, bit 16+t <- column 2t+1 inside each 32-column group
    const uint32_t mw = bits[(size_t)row * bits_ld + (k >> 5)] >> ((k & 31) >> 1);
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < OUT; ++o) s = fmaf(d[o], w[o][e], s);
      f[e] = ((mw >> ((e >> 1) + 16 * (e & 1))) & 1u) ? s : 0.f;
    }
    *reinterpret_cast<uint4*>(dY + (size_t)row * ldy + k) =
        make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]), tc::pack_bf16(f[6], f[7]));
  }
}

// bit c of row = (X[row,c] > 0)   (diagnostics: builds the mask the forward epilogue would have written)
static __global__ void k_make_bits(int rows, int width, const float* __restrict__ X, uint32_t* __restrict__ bits) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int words = width / 32;
  if (i >= rows * words) return;
  int r = i / words, wd = i - r * words;
  uint32_t m = 0;
  for (int e = 0; e < 32; ++e) m |= (X[(size_t)r * width + wd * 32 + e] > 0.f ? 1u : 0u) << ((e >> 1) + 16 * (e & 1));
  bits[i] = m;
}

// dW[o,k] += sum_row dl[row,o] X[row,k] ; db[o] += sum_row dl[row,o]   (block = (width/8, 8); smem reduce over y;
// one fp32 atomic per (o,k) per block)
template <int OUT>
__global__ void k_thin_dw(int n, int width, const float* __restrict__ dl, const bf16* __restrict__ X, int ld,
                          float* __restrict__ dW, int ldw, float* __restrict__ db, int rows_per_block) {
  pdl_wait();
  extern __shared__ float red[];      // [blockDim.y][OUT*width + OUT]
  const int kq = threadIdx.x * 8;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
  float acc[OUT][8];
  float bacc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) { bacc[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[o][e] = 0.f; }
  constexpr int R = 4;
  for (int rowb = r0 + threadIdx.y; rowb < r1; rowb += blockDim.y * R) {
    uint4 raw[R];
    float4 d4[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int row = rowb + i * blockDim.y;
      const bool ok = row < r1;
      raw[i] = ok ? *reinterpret_cast<const uint4*>(X + (size_t)row * ld + kq) : make_uint4(0u, 0u, 0u, 0u);
      d4[i] = ok ? *reinterpret_cast<const float4*>(dl + (size_t)row * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float d[3] = {d4[i].x, d4[i].y, d4[i].z};
      const uint32_t w4[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int o = 0; o < OUT; ++o) bacc[o] += d[o];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = __uint_as_float(((w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu) << 16);
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o][e] = fmaf(d[o], xv, acc[o][e]);
      }
    }
  }
  const int stride = OUT * width + OUT;
  float* mine = red + threadIdx.y * stride;
#pragma unroll
  for (int o = 0; o < OUT; ++o) {
#pragma unroll
    for (int e = 0; e < 8; ++e) mine[o * width + kq + e] = acc[o][e];
    if (threadIdx.x == 0) mine[OUT * width + o] = bacc[o];
  }
  __syncthreads();
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = t; i < stride; i += blockDim.x * blockDim.y) {
    float ssum = 0.f;
    for (int y = 0; y < (int)blockDim.y; ++y) ssum += red[y * stride + i];
    if (i < OUT * width) atomicAdd(&dW[(size_t)(i / width) * ldw + (i % width)], ssum);
    else atomicAdd(&db[i - OUT * width], ssum);
  }
}

static __global__ void k_bf16_to_f32(long long n, const bf16* __restrict__ in, float* __restrict__ out) {
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

// table-driven (un)packing: one launch for every layer of both networks
// mode 0: f32 pad, 1: bf16, 2: bf16 transposed, 3: bf16 hi rows [0,8) / lo rows [8,16); wcols > 0: only columns < wcols are written
struct PackEntry { const float* src; void* dst; int rows, cols, prow, pcol, mode, src_ld, col_off, wcols; };
constexpr int kMaxPack = 40;
// The prologue of a step in ONE launch: rows [0, n) of the grid pack the weights (table-driven); when `fused` the three rows
// after them compute the homographies H = exp(A(h)) of all patches, the mask head's colour-class table, and zero the
// accumulators of the step (gradient twins, per-patch Jacobians, class-table scratch, loss sums).
struct PackTable {
  PackEntry e[kMaxPack]; int n;
  int fused;
  const float* warp; int n_patches; float* Hm;                                            // row n
  const float* mW0; const float* mb0; const float* embed; int mk_in, mk_out, edim, k_uv; bf16* mWk; int mldk;   // row n+1 (mk_out == 0: none)
  void* zptr[5]; unsigned long long zbytes[5];                                            // row n+2 (8-byte multiples)
};
static __global__ void k_pack_table(const __grid_constant__ PackTable t) {
  pdl_wait();
  if ((int)blockIdx.y >= t.n) {
    const int role = (int)blockIdx.y - t.n;
    if (role == 0) {
      __shared__ double X[9], T[9], R[9];
      for (int b = blockIdx.x; b < t.n_patches; b += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) sl3_generator(t.warp + 8 * b, X);
        __syncthreads();
        expm_coop<3>(X, T, R, threadIdx.x);
        if (threadIdx.x < 9) t.Hm[9 * b + threadIdx.x] = (float)R[threadIdx.x];
      }
    } else if (role == 1) {
      const int nvb = (8 * t.mk_out * 32 + (int)blockDim.x - 1) / (int)blockDim.x;
      for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x)
        mask_table_block(vb, t.mW0, t.mb0, t.embed, t.mk_in, t.mk_out, t.edim, t.k_uv, t.mWk, t.mldk);
    } else {
      for (int z = 0; z < 5; ++z) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(t.zptr[z]);
        const unsigned long long n8 = t.zbytes[z] / 8;
        for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (unsigned long long)gridDim.x * blockDim.x)
          p[i] = 0ull;
      }
    }
    return;
  }
  const PackEntry& E = t.e[blockIdx.y];
  const int tot = E.prow * E.pcol;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
    int pr = i / E.pcol, pc = i - pr * E.pcol;
    if (E.wcols > 0 && pc >= E.wcols) continue;
    int r = E.mode == 2 ? pc : pr, c = E.mode == 2 ? pr : pc;
    if (E.mode == 3) r = pr & 7;
    float v = (r < E.rows && c < E.cols) ? E.src[(size_t)r * E.src_ld + E.col_off + c] : 0.f;
    if (E.mode == 3 && pr >= 8) v -= __bfloat162float(__float2bfloat16(v));     // the part the hi row misses
    if (E.mode == 0) reinterpret_cast<float*>(E.dst)[i] = v;
    else reinterpret_cast<bf16*>(E.dst)[i] = __float2bfloat16(v);
  }
}
// The tail of the backward pass in ONE launch: every gradient from its padded fp32 twin to the caller's tensor (mode 0), the
// mask head's layer-0 weight / bias gradients straight from the class-table scratch (modes 1 / 2, without a
// round trip through the padded twin), and — row t.n of the grid — the sl(3) adjoint of every owned
// patch plus the zero rows of the patches other ranks own.
struct UnpackEntry { const float* src; float* dst; int rows, cols, pcol; int mode; const float* embed; int edim, k_uv; };
struct UnpackTable {
  UnpackEntry e[kMaxPack]; int n;
  const float* warp; const double* G; float* g_warp; int patch_offset, n_local, n_global;     // sl(3) row (g_warp == nullptr: none)
};
static __global__ void k_unpack_table(const __grid_constant__ UnpackTable t) {
  pdl_wait();
  if ((int)blockIdx.y == t.n) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < t.n_global * 8; b += gridDim.x * blockDim.x) {
      const int p = b >> 3;
      if (p < t.patch_offset || p >= t.patch_offset + t.n_local) t.g_warp[b] = 0.f;
    }
    for (int bl = blockIdx.x; bl < t.n_local; bl += gridDim.x) sl3_backward_block(t.warp, t.G, t.patch_offset, bl, t.g_warp);
    return;
  }
  const UnpackEntry& E = t.e[blockIdx.y];
  const int tot = E.rows * E.cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += gridDim.x * blockDim.x) {
    int r = i / E.cols, c = i - r * E.cols;
    if (E.mode == 0) {
      E.dst[i] = E.src[(size_t)r * E.pcol + c];
    } else {
      // src = X [k_out, 64] = dY0^T [uv | onehot | onehot]: uv columns -> dW0[:, 3E : 3E + k_uv]; per-class sums S[c] = X[r][k_uv + c]
      const float* x = E.src + (size_t)(E.mode == 1 ? r : c) * 64;
      float s0 = 0.f, s1 = 0.f;
      const int ch = E.mode == 1 && c < 3 * E.edim ? c / E.edim : 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = x[E.k_uv + k];
        if ((k >> ch) & 1) s1 += v; else s0 += v;
      }
      if (E.mode == 2) E.dst[i] = s0 + s1;                                     // db0[j] = sum over the classes
      else if (c >= 3 * E.edim) E.dst[i] = x[c - 3 * E.edim];                  // uv columns
      else E.dst[i] = s0 * E.embed[c - ch * E.edim] + s1 * E.embed[E.edim + c - ch * E.edim];
    }
  }
}

// ------------------------------------------------------------------------------------------------ state
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct BfLayer {
  int k_in, k_out;       // true dims
  int kp, np;            // padded: kp = round64(k_in), np = round64(k_out) (thin layers: unused)
  bool thin;             // k_out <= 4: SIMT output layer
  bf16* Wk = nullptr;    // [np, kp]   forward B operand (K-major)
  bf16* Wt = nullptr;    // [kp, np]   dX B operand (K-major over out features)
  CUtensorMap tmWk, tmWt;
  CUtensorMap tmWk128, tmWt128;   // 128-row boxes: the weight chunks of the fused chain kernel
  CUtensorMap tmWk64, tmWt64;     // 64-row boxes: one CTA's half of a weight chunk (CTA-pair chain kernel)
};

struct BfChain {
  int n = 0;
  BfLayer L[MARF_MAX_LAYERS];
  bf16* act[MARF_MAX_LAYERS + 1] = {};    // act[l]: input of layer l [chunk, ld[l]]
  int ld[MARF_MAX_LAYERS + 1] = {};
  CUtensorMap tmAct128[MARF_MAX_LAYERS + 1];   // box {64,128}: GEMM A loads / epilogue stores / mask loads
  CUtensorMap tmAct64[MARF_MAX_LAYERS + 1];    // box {64,64}: dW loads
  uint32_t* bits[MARF_MAX_LAYERS + 1] = {};  // bits[l]: 1-bit ReLU mask of act[l] (l >= 1), [chunk, ld[l]/32] words
  bf16* dY[MARF_MAX_LAYERS] = {};         // dY[l]: gradient wrt the output of layer l [chunk, np(l)] (kept for the dW pass)
  CUtensorMap tmDY128[MARF_MAX_LAYERS], tmDY64[MARF_MAX_LAYERS];
  uint32_t* flags_fwd[MARF_MAX_LAYERS + 1] = {};   // flags_fwd[l]: per-tile completion counters of act[l] (written by layer l-1)
  uint32_t* flags_bwd[MARF_MAX_LAYERS] = {};       // flags_bwd[l]: per-tile completion counters of dY[l] (written by dX of layer l+1)
  float* logits = nullptr;                // [chunk,4] fp32
  float* dlogits = nullptr;               // [chunk,4] fp32
  bf16* dl16 = nullptr;                   // [chunk,8] bf16 copy of dlogits: A operand of the output layer's dW / dX (64-wide TMA boxes, OOB zero fill)
  CUtensorMap tmDL64, tmDL128;
  bf16* Wlast_t = nullptr;                // [k_in(last), 64] bf16: W_last^T zero padded (B operand of the output layer's dX GEMM)
  CUtensorMap tmWlast_t, tmWlast128, tmWlast64;
  bf16* Wout16 = nullptr;                 // [16, k_in(last)] bf16: rows 0..7 hi, 8..15 lo halves of W_last (fused chain output layer)
  CUtensorMap tmWout16, tmWout8;
  bool fused = false;                     // shape served by k_tc_chain: 64 -> 256 x4 -> (<=4)
  Chain* f32 = nullptr;                   // padded fp32 twin (gradient accumulators, bias)
  bool need_dx0 = false;
  int col_off0 = 0;                       // class-table mode: layer 0 uses columns [col_off0, col_off0 + k_in) of W0
  float* dW0x = nullptr;                  // [256, 64] fp32: dY0^T [uv | onehot | onehot] (class-table mode)
  float* zero_bias = nullptr;             // [256] zeros: the class table carries b0
};

struct Bf16State {
  EncodeTiledFn encode = nullptr;
  BfChain img, msk;
  float* dX0 = nullptr;                   // [chunk, 64] fp32
  int num_sms = 148;
  bool table_done = false;                // the class table of this step was written by the fused prologue launch
  bool accs_zeroed = false;               // the prologue launch of the forward pass already zeroed the backward accumulators
  uint32_t* ready = nullptr;              // k_tc_bwd: [2 chains][4 units][chunk / 128] per-tile hand-over flags (epoch values, never reset)
  uint32_t epoch = 0;                     // k_tc_bwd launches of this handle so far
  uint32_t* flags_all = nullptr;          // every per-tile flag array of both chains, zeroed before each chained launch
  size_t flags_words = 0, flags_used = 0;
};

static int make_tmap(marf_handle* h, Bf16State* S, CUtensorMap* m, void* base, int rows, int cols_ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols_ld, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols_ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = S->encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MARF_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  return MARF_OK;
}

static int round64(int a) { return (a + 63) / 64 * 64; }

static int build_bf_chain(marf_handle* h, Bf16State* S, BfChain& B, Chain& F, bool need_dx0, int class_cols = 0) {
  B.n = F.n;
  B.f32 = &F;
  B.need_dx0 = need_dx0;
  B.col_off0 = class_cols;
  if (class_cols > 0) {
    B.dW0x = (float*)ws_alloc(h, 256 * 64 * sizeof(float));
    B.zero_bias = (float*)ws_alloc(h, 256 * sizeof(float));
    if (!B.dW0x || !B.zero_bias) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (class table)");
    if (F.k_out[0] != 256 || F.k_in[0] - class_cols + 16 > 64)
      return fail(h, MARF_ERR_UNSUPPORTED, "bf16 mask head: layer 0 must be 256 wide with at most 48 positional inputs");
  }
  for (int l = 0; l < F.n; ++l) {
    BfLayer& L = B.L[l];
    L.k_in = l == 0 ? F.k_in[l] - class_cols : F.k_in[l];
    L.k_out = F.k_out[l];
    L.thin = L.k_out <= 4;
    L.kp = round64(L.k_in);
    L.np = round64(L.k_out);
    B.ld[l] = L.kp;
    if (!L.thin) {
      if (l == F.n - 1) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: the last layer must be the 3-/1-wide output layer");
      if (L.k_out != 256 && L.k_out != 512)
        return fail(h, MARF_ERR_UNSUPPORTED, "bf16: hidden widths 256 and 512 are implemented (got " + std::to_string(L.k_out) + ")");
      if (L.kp > 512) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: layer input wider than 512 does not fit the resident-weight tile");
      if (l >= 1 && L.kp != 256 && L.kp != 512 && !(l == 0))
        return fail(h, MARF_ERR_UNSUPPORTED, "bf16: hidden layer inputs must be 256 or 512 wide");
      L.Wk = (bf16*)ws_alloc(h, (size_t)L.np * L.kp * 2);
      L.Wt = (bf16*)ws_alloc(h, (size_t)L.kp * L.np * 2);
      if (!L.Wk || !L.Wt) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed");
      int nt = (L.kp > 256 && L.np > 128) ? 128 : std::min(L.np, 256);   // forward N tile rows per TMA box
      int rc = make_tmap(h, S, &L.tmWk, L.Wk, L.np, L.kp, nt);
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWt, L.Wt, L.kp, L.np, std::min(L.kp, 256));
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWk128, L.Wk, L.np, L.kp, std::min(L.np, 128));
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWt128, L.Wt, L.kp, L.np, std::min(L.kp, 128));
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWk64, L.Wk, L.np, L.kp, 64);
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWt64, L.Wt, L.kp, L.np, 64);
      if (rc) return rc;
    } else if (l != F.n - 1) {
      return fail(h, MARF_ERR_UNSUPPORTED, "bf16: thin hidden layers are not supported");
    } else if (L.k_in != 256 && L.k_in != 512) {
      return fail(h, MARF_ERR_UNSUPPORTED, "bf16: the output layer must read 256 or 512 features");
    }
  }
  B.ld[F.n] = 4;
  for (int l = 0; l < F.n; ++l) {
    B.act[l] = (bf16*)ws_alloc(h, (size_t)h->chunk * B.ld[l] * 2);
    if (!B.act[l]) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (activations)");
    int rc = make_tmap(h, S, &B.tmAct128[l], B.act[l], h->chunk, B.ld[l], 128);
    if (rc) return rc;
    rc = make_tmap(h, S, &B.tmAct64[l], B.act[l], h->chunk, B.ld[l], 64);
    if (rc) return rc;
    if (l >= 1) {
      B.bits[l] = (uint32_t*)ws_alloc(h, (size_t)h->chunk * (B.ld[l] / 32) * 4);
      B.flags_fwd[l] = S->flags_all + S->flags_used;
      S->flags_used += h->chunk / 128 + 1;
      if (!B.bits[l] || S->flags_used > S->flags_words) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (mask bits)");
    }
    if (l < F.n - 1) {
      B.flags_bwd[l] = S->flags_all + S->flags_used;
      S->flags_used += h->chunk / 128 + 1;
      if (S->flags_used > S->flags_words) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (flags)");
    }
    if (l < F.n - 1) {
      B.dY[l] = (bf16*)ws_alloc(h, (size_t)h->chunk * B.L[l].np * 2);
      if (!B.dY[l]) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (dY)");
      rc = make_tmap(h, S, &B.tmDY128[l], B.dY[l], h->chunk, B.L[l].np, 128);
      if (rc) return rc;
      rc = make_tmap(h, S, &B.tmDY64[l], B.dY[l], h->chunk, B.L[l].np, 64);
      if (rc) return rc;
    }
  }
  B.logits = (float*)ws_alloc(h, (size_t)h->chunk * 4 * sizeof(float));
  B.dlogits = (float*)ws_alloc(h, (size_t)h->chunk * 4 * sizeof(float));
  // dlogits as an MMA operand: 8 bf16 per row in HBM; the 64-column TMA boxes zero-fill the other 56 (out of bounds)
  B.dl16 = (bf16*)ws_alloc(h, (size_t)h->chunk * 8 * 2);
  if (!B.logits || !B.dlogits || !B.dl16) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (logits)");
  int rc = make_tmap(h, S, &B.tmDL64, B.dl16, h->chunk, 8, 64);
  if (rc) return rc;
  rc = make_tmap(h, S, &B.tmDL128, B.dl16, h->chunk, 8, 128);
  if (rc) return rc;
  const int kl = B.L[F.n - 1].k_in;
  if (kl == 256) {
    B.Wlast_t = (bf16*)ws_alloc(h, (size_t)kl * 64 * 2);
    if (!B.Wlast_t) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (output layer)");
    rc = make_tmap(h, S, &B.tmWlast_t, B.Wlast_t, kl, 64, 256);
    if (rc) return rc;
    rc = make_tmap(h, S, &B.tmWlast128, B.Wlast_t, kl, 64, 128);
    if (rc) return rc;
    B.Wout16 = (bf16*)ws_alloc(h, (size_t)16 * kl * 2);
    if (!B.Wout16) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (output layer)");
    rc = make_tmap(h, S, &B.tmWout16, B.Wout16, 16, kl, 16);
    if (rc) return rc;
    rc = make_tmap(h, S, &B.tmWout8, B.Wout16, 16, kl, 8);
    if (rc) return rc;
    rc = make_tmap(h, S, &B.tmWlast64, B.Wlast_t, kl, 64, 64);
    if (rc) return rc;
    bool ok = B.n - 1 == tc::kChUnits && B.L[0].kp == 64 && B.L[B.n - 1].k_out <= 4;
    for (int l = 0; l < B.n - 1; ++l) ok = ok && B.L[l].np == 256 && (l == 0 || B.L[l].kp == 256);
    B.fused = ok && getenv("MARF_NO_FUSE") == nullptr;
  }
  return rc;
}

static int set_tc_attrs(marf_handle* h) {
  const int big = 232448;
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_PLAIN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_FWD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_DX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_FWD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_DX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_bwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_FWD, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_chain<tc::CH_DX, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(k_thin_dw<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (3 * 512 + 3) * 4));
  BF_TRY(h, cudaFuncSetAttribute(k_thin_dw<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1 * 512 + 1) * 4));
  return MARF_OK;
}

int bf16_create(marf_handle* h) {
  Bf16State* S = new Bf16State();
  h->bf16 = S;
  if (h->cfg.skip_mask) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: arch.skip is only implemented for precision=fp32");
  if (h->geo.d_in > 64) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: posenc wider than 64 inputs (L_2D > 15)");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
    return fail(h, MARF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  S->encode = (EncodeTiledFn)fn;
  cudaDeviceProp prop;
  BF_TRY(h, cudaGetDeviceProperties(&prop, h->cfg.device));
  S->num_sms = prop.multiProcessorCount;
  S->flags_words = (size_t)(h->chunk / 128 + 1) * 4 * MARF_MAX_LAYERS;
  S->flags_all = (uint32_t*)ws_alloc(h, S->flags_words * 4);
  S->ready = (uint32_t*)ws_alloc(h, (size_t)2 * tc::kChUnits * (h->chunk / 128 + 1) * 4);
  if (!S->flags_all || !S->ready) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (flags)");
  int rc = build_bf_chain(h, S, S->img, h->img, true);
  if (rc) return rc;
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT) {
    rc = build_bf_chain(h, S, S->msk, h->msk, false, 3 * h->cfg.mask_embed_dim);
    if (rc) return rc;
  }
  S->dX0 = (float*)ws_alloc(h, (size_t)h->chunk * 64 * sizeof(float));
  if (!S->dX0) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (dX0)");
  return set_tc_attrs(h);
}

void bf16_destroy(marf_handle* h) {
  delete h->bf16;
  h->bf16 = nullptr;
}

bool bf16_supported(const marf_handle* h, const marf_step_io*, std::string* why) {
  if (!h->bf16) { if (why) *why = "handle was not created with precision=bf16"; return false; }
  return true;
}

// ------------------------------------------------------------------------------------------------ launches
// marf_profile: an event pair around the launches of one kernel class (scope object; no-op unless profiling is on)
struct ProfScope {
  marf_handle* h; cudaStream_t st; int cls; cudaEvent_t e1 = nullptr;
  static cudaEvent_t get(marf_handle* h) {
    if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  ProfScope(marf_handle* h_, cudaStream_t st_, int cls_) : h(h_), st(st_), cls(cls_) {
    if (!h->profiling) return;
    cudaEvent_t e0 = get(h);
    e1 = get(h);
    cudaEventRecord(e0, st);
    h->prof_ev[cls].push_back({e0, e1});
  }
  ~ProfScope() { if (e1) cudaEventRecord(e1, st); }
};

// diagnostics (MARF_TC_TRACE=<n>): dump per-tile clock64 stamps of the first CTA of every job of the n-th traced launch
struct TraceCtx { long long* dev = nullptr; int iters = 0; int njobs = 0; };
static TraceCtx trace_begin(tc::GemmJobs& jobs) {
  TraceCtx t;
  static int calls = 0;
  const char* e = getenv("MARF_TC_TRACE");
  if (!e || ++calls != atoi(e)) return t;
  t.njobs = jobs.n;
  for (int i = 0; i < jobs.n; ++i) t.iters = std::max(t.iters, (jobs.j[i].p.n_tiles + jobs.j[i].cta_count - 1) / jobs.j[i].cta_count);
  cudaMalloc(&t.dev, (size_t)jobs.n * t.iters * 16 * sizeof(long long));
  cudaMemset(t.dev, 0, (size_t)jobs.n * t.iters * 16 * sizeof(long long));
  for (int i = 0; i < jobs.n; ++i) jobs.j[i].p.trace = t.dev + (size_t)i * t.iters * 16;
  return t;
}
static void trace_end(const TraceCtx& t, cudaStream_t st, const char* what) {
  if (!t.dev) return;
  cudaStreamSynchronize(st);
  std::vector<long long> tr((size_t)t.njobs * t.iters * 16);
  cudaMemcpy(tr.data(), t.dev, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  for (int i = 0; i < t.njobs; ++i) {
    long long t00 = tr[(size_t)i * t.iters * 16 + 1];
    fprintf(stderr, "trace %s job %d: iter | prod mma_start mma_commit | g0 acc_full ld0 st0 ld1 st1 | g1 acc_full ld0 st0 ld1 st1\n", what, i);
    for (int it = 0; it < t.iters; ++it) {
      fprintf(stderr, "%3d |", it);
      for (int k = 0; k < 15; ++k) {
        long long v = tr[((size_t)i * t.iters + it) * 16 + k];
        fprintf(stderr, " %7lld", v ? v - t00 : -1);
      }
      fprintf(stderr, "\n");
    }
  }
  cudaFree(t.dev);
}

// distribute the launch's CTAs over the jobs in proportion to `weight` (>= 1 CTA each), at most n_tiles per job
static void assign_ctas(tc::GemmJobs& jobs, const int* weight, int num_sms) {
  int wsum = 0;
  for (int i = 0; i < jobs.n; ++i) wsum += weight[i];
  int begin = 0, left = num_sms;
  for (int i = 0; i < jobs.n; ++i) {
    int c = std::max(1, (int)((long long)num_sms * weight[i] / wsum));
    c = std::min(c, std::min(jobs.j[i].p.n_tiles, left - (jobs.n - 1 - i)));
    jobs.j[i].cta_begin = begin;
    jobs.j[i].cta_count = std::max(1, c);
    begin += jobs.j[i].cta_count;
    left -= jobs.j[i].cta_count;
  }
}

static tc::GemmJob fwd_job(BfChain& B, int l, int rows, int n_tile, int n0) {
  BfLayer& L = B.L[l];
  tc::GemmJob J{};
  J.tmA = B.tmAct128[l];
  J.tmW = L.tmWk;
  J.tmOut = B.tmAct128[l + 1];
  J.p.n_tiles = rows / 128;
  J.p.k_chunks = L.kp / 64;
  J.p.bias = B.f32->bp[l];
  if (l == 0 && B.col_off0 > 0) J.p.bias = B.zero_bias;     // b0 is part of the class table inside Wk
  J.p.bits_out = B.bits[l + 1];
  J.p.bits_ld = B.ld[l + 1] / 32;
  J.p.reverse = getenv("MARF_CHAIN") ? 0 : (l & 1);   // start where the previous launch just finished (still in L2)
  J.p.load_policy = tc::kEvictFirst;           // inputs are not read again before the backward pass
  J.p.prefetch_ahead = getenv("MARF_PREFETCH") ? atoi(getenv("MARF_PREFETCH")) : 0;
  J.p.store_policy = tc::kEvictLast;           // outputs are the next launch's inputs
  J.n0 = n0;
  (void)n_tile;
  return J;
}

// forward of every tensor-core layer of the given chains.  Layers whose input is wider than 256 (mask head layer 0)
// run first as their own launch (N split in two 128-column jobs); all other layers of all chains share ONE launch
// in which layer l+1 consumes layer l's tiles through per-tile flags (L2-resident hand-over).
static int launch_forward_all(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows) {
  Bf16State* S = h->bf16;
  // a layer whose input is wider than 256 (width-512 networks): N split into 128-column jobs, [128, K] weight tiles resident
  auto launch_wide = [&](BfChain& B, int l) -> int {
    BfLayer& L = B.L[l];
    tc::GemmJobs jobs{};
    int w[tc::kMaxGemmJobs];
    if (L.np / 128 > tc::kMaxGemmJobs) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 forward: layer too wide");
    for (int t = 0; t < L.np / 128; ++t) { jobs.j[jobs.n] = fwd_job(B, l, rows, 128, t * 128); w[jobs.n++] = 1; }
    assign_ctas(jobs, w, S->num_sms);
    int smem = tc::gemm_smem(128, L.kp / 64, true).total + 1024;
    int grid = jobs.j[jobs.n - 1].cta_begin + jobs.j[jobs.n - 1].cta_count;
    launch_k(tc::k_tc_gemm<128, tc::EPI_BIAS_RELU>, grid, tc::kThreads, smem, st, jobs);
    BF_LAUNCH(h);
    return MARF_OK;
  };
  // Chained mode (MARF_CHAIN=1, experimental): every remaining layer of every chain in one launch, layer l+1 consuming
  // layer l's tiles through per-tile flags.  Default: one launch per depth level (the same-depth layers of all chains
  // share the SMs); measured faster on B200 because each launch then streams at the full HBM rate.
  static const bool chain_mode = getenv("MARF_CHAIN") != nullptr;
  int max_depth = 0;
  for (int ci = 0; ci < n_chains; ++ci) max_depth = std::max(max_depth, chains[ci]->n - 1);
  if (chain_mode) BF_TRY(h, cudaMemsetAsync(S->flags_all, 0, S->flags_used * 4, st));
  // default: one layer per launch, chain after chain: a 110 MB output stays L2-resident for the next launch, which
  // walks the tiles in the opposite direction (measured: faster than sharing a launch between the two chains)
  for (int pass = 0; pass < (chain_mode ? 1 : max_depth * n_chains); ++pass) {
    const int only_chain = pass / max_depth, level = pass % max_depth;
    tc::GemmJobs jobs{};
    int w[tc::kMaxGemmJobs];
    int max_kc = 1;
    for (int ci = 0; ci < n_chains; ++ci) {
      BfChain& B = *chains[ci];
      for (int l = 0; l < B.n - 1; ++l) {
        if (!chain_mode && (l != level || ci != only_chain)) continue;
        BfLayer& L = B.L[l];
        if (L.kp > 256 && L.np > 128) {
          if (chain_mode) return fail(h, MARF_ERR_UNSUPPORTED, "MARF_CHAIN does not serve layers wider than 256");
          int rc = launch_wide(B, l);               // (one layer per pass here: layer order is preserved)
          if (rc) return rc;
          continue;
        }
        if (L.np % 256) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 forward: hidden width must be a multiple of 256");
        for (int t = 0; t < L.np / 256; ++t) {      // (width 512: two 256-column jobs)
          if (jobs.n >= tc::kMaxGemmJobs) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 forward: too many layers");
          tc::GemmJob J = fwd_job(B, l, rows, 256, t * 256);
          if (chain_mode && L.np == 256) {
            const bool chained_in = l >= 1 && !(B.L[l - 1].kp > 256 && B.L[l - 1].np > 128);
            J.flags_in = chained_in ? B.flags_fwd[l] : nullptr;
            J.in_target = 2u;
            J.flags_out = (l + 1 < B.n - 1) ? B.flags_fwd[l + 1] : nullptr;
          }
          w[jobs.n] = 1;                            // the epilogue (same for every layer) bounds a tile, not K
          max_kc = std::max(max_kc, L.kp / 64);
          jobs.j[jobs.n++] = J;
        }
      }
    }
    if (jobs.n == 0) continue;
    assign_ctas(jobs, w, S->num_sms);
    int smem = tc::gemm_smem(256, max_kc, true).total + 1024;
    int grid = jobs.j[jobs.n - 1].cta_begin + jobs.j[jobs.n - 1].cta_count;
    TraceCtx tr = trace_begin(jobs);
    launch_k(tc::k_tc_gemm<256, tc::EPI_BIAS_RELU>, grid, tc::kThreads, smem, st, jobs);
    BF_LAUNCH(h);
    trace_end(tr, st, "fwd");
  }
  return MARF_OK;
}

// dX0 of the chains that need the gradient w.r.t. their input: the epilogue is the backward of the encoding prologue
static int launch_dx0(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows, PxRange rg) {
  Bf16State* S = h->bf16;
  ProfScope prof(h, st, MARF_PROF_DX0);
  for (int ci = 0; ci < n_chains; ++ci) {
    BfChain& B = *chains[ci];
    if (!B.need_dx0) continue;
    BfLayer& L = B.L[0];
    if (L.kp != 64) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX0: encoded input must fit 64 columns");
    tc::GemmJobs one{};
    one.n = 1;
    tc::GemmJob& J = one.j[0];
    J.tmA = B.tmDY128[0];
    J.tmW = L.tmWt;
    J.tmOut = B.tmDY128[0];
    J.p.n_tiles = rows / 128;
    J.p.k_chunks = L.np / 64;
    J.p.load_policy = tc::kEvictNormal;
    J.p.store_policy = tc::kEvictNormal;
    J.p.geo = h->geo;
    J.p.rg = rg;
    J.p.Hm = h->Hm;
    J.p.G = h->G;
    J.cta_begin = 0;
    J.cta_count = std::min(J.p.n_tiles, S->num_sms);
    int smem = tc::gemm_smem(64, J.p.k_chunks, false).total + 1024;
    // the epilogue is the backward of the encoding prologue: per-pixel (g_u, g_v) -> dq -> per-patch G
    if (h->geo.L == 8) launch_k(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 8>, J.cta_count, tc::kThreads, smem, st, one);
    else if (h->geo.L == 10) launch_k(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 10>, J.cta_count, tc::kThreads, smem, st, one);
    else if (h->geo.L == 4) launch_k(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 4>, J.cta_count, tc::kThreads, smem, st, one);
    else launch_k(tc::k_tc_gemm<64, tc::EPI_WARP_GRAD, 0>, J.cta_count, tc::kThreads, smem, st, one);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

// dX of every tensor-core layer l >= 1 of the given chains in ONE launch (dY[l-1] = (dY[l] W_l) * relu_mask(act[l])),
// chained through per-tile flags; then dX0 of the chains that need the gradient w.r.t. their input.
static int launch_dx_all(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows, PxRange rg) {
  Bf16State* S = h->bf16;
  static const bool chain_mode = getenv("MARF_CHAIN") != nullptr;
  int max_depth = 0;
  for (int ci = 0; ci < n_chains; ++ci) max_depth = std::max(max_depth, chains[ci]->n - 1);
  if (chain_mode) BF_TRY(h, cudaMemsetAsync(S->flags_all, 0, S->flags_used * 4, st));
  for (int pass = 0; pass < (chain_mode ? 1 : max_depth * n_chains); ++pass) {
    const int only_chain = pass / max_depth, level = max_depth - 1 - pass % max_depth;
    tc::GemmJobs jobs{};
    int w[tc::kMaxGemmJobs];
    bool wide = false;
    int max_kc = 4;
    for (int ci = 0; ci < n_chains; ++ci) {
      BfChain& B = *chains[ci];
      for (int l = B.n - 2; l >= 1; --l) {
        if (!chain_mode && (l != level || ci != only_chain)) continue;
        BfLayer& L = B.L[l];
        // 256 -> 256: one job with the whole W^T resident; wider layers: 128-column jobs (W^T tile [128, np] resident)
        wide = L.kp != 256 || L.np != 256;
        if (wide && (L.kp % 128 || L.np > 512)) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX: layer shape not supported");
        const int n_tile = wide ? 128 : 256;
        max_kc = std::max(max_kc, L.np / 64);
        for (int t = 0; t < L.kp / n_tile; ++t) {
          if (jobs.n >= tc::kMaxGemmJobs) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX: too many layers");
          tc::GemmJob J{};
          J.tmA = B.tmDY128[l];
          J.tmW = wide ? L.tmWt128 : L.tmWt;
          J.tmOut = B.tmDY128[l - 1];
          J.p.n_tiles = rows / 128;
          J.p.k_chunks = L.np / 64;
          J.p.bits_in = B.bits[l];
          J.p.bits_ld = B.ld[l] / 32;
          J.p.reverse = chain_mode ? 0 : (l & 1);    // start where the previous launch just finished (still in L2)
          J.p.load_policy = tc::kEvictFirst;       // (dY[l] is read again by the dW pass, but long after L2 has turned over)
          J.p.store_policy = tc::kEvictLast;
          J.n0 = t * n_tile;
          if (chain_mode && !wide) {
            J.flags_in = l < B.n - 2 ? B.flags_bwd[l] : nullptr;   // dY[n-2] comes from the output-layer kernel (complete)
            J.in_target = 2u;
            J.flags_out = l - 1 >= 1 ? B.flags_bwd[l - 1] : nullptr;
          }
          w[jobs.n] = 1;
          jobs.j[jobs.n++] = J;
        }
      }
    }
    if (jobs.n == 0) continue;
    assign_ctas(jobs, w, S->num_sms);
    int grid = jobs.j[jobs.n - 1].cta_begin + jobs.j[jobs.n - 1].cta_count;
    if (wide) launch_k(tc::k_tc_gemm<128, tc::EPI_RELU_MASK>, grid, tc::kThreads, tc::gemm_smem(128, max_kc, true).total + 1024, st, jobs);
    else launch_k(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, grid, tc::kThreads, tc::gemm_smem(256, 4, true).total + 1024, st, jobs);
    BF_LAUNCH(h);
  }
  return launch_dx0(h, st, chains, n_chains, rows, rg);
}

// all dW / db of the tensor-core layers of the given chains in ONE launch; the CTAs are split over the jobs by the bytes
// each job streams
static int launch_dw_all(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows, long long row_first) {
  Bf16State* S = h->bf16;
  (void)row_first;
  std::vector<tc::DwJob> all;
  std::vector<int> weight;
  auto add = [&](const CUtensorMap& tmDY, const CUtensorMap& tmX, int n_tile, int m0, int m_layer, int n_valid, int n0, int ld_w,
                 int do_bias, float* dW, float* db) {
    tc::DwJob J{};
    J.tmDY = tmDY; J.tmX = tmX; J.rows = rows; J.n_tile = n_tile; J.m0 = m0;
    J.m_halves = (std::min(m_layer - m0, 256) + 127) / 128; J.m_valid = m_layer;
    J.n_valid = n_valid; J.n0 = n0; J.ld_w = ld_w; J.do_bias = do_bias; J.dW = dW; J.db = db;
    all.push_back(J);
    weight.push_back(J.m_halves * 128 + n_tile);         // bytes per pixel row ~ (dY columns + X columns)
  };
  for (int ci = 0; ci < n_chains; ++ci) {
    BfChain& B = *chains[ci];
    Chain& F = *B.f32;
    if (B.L[B.n - 1].k_in == 256) {
      // output layer (3-/1-wide): dW = dlogits^T X_last through the same kernel (dlogits as a zero-padded bf16 tile)
      const int l = B.n - 1;
      add(B.tmDL64, B.tmAct64[l], 256, 0, B.L[l].k_out, B.L[l].k_in, 0, F.ld_in[l], 1, F.gWp[l], F.gbp[l]);
    }
    for (int l = 0; l < B.n - 1; ++l) {
      BfLayer& L = B.L[l];
      const int n_tile = L.kp >= 256 ? 256 : 64;
      if (n_tile == 64 && L.kp != 64) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dW: input width must be 64 or >= 256");
      const int n_tiles_n = (L.kp + n_tile - 1) / n_tile;
      for (int m0 = 0; m0 < L.k_out; m0 += 256)          // (outputs wider than 256: one job per 256 output features)
        for (int t = 0; t < n_tiles_n; ++t) {
          if (l == 0 && B.col_off0 > 0)
            // class-table mode: the whole [256, 64] tile (uv columns and per-class sums) goes to a scratch that
            // the tail launch (k_unpack_table modes 1 / 2) turns into dW0 / db0
            add(B.tmDY64[l], B.tmAct64[l], n_tile, m0, L.k_out, 64, t * n_tile, 64, 0, B.dW0x, F.gbp[l]);
          else
            add(B.tmDY64[l], B.tmAct64[l], n_tile, m0, L.k_out, L.k_in, t * n_tile, F.ld_in[l], t == 0, F.gWp[l], F.gbp[l]);
        }
    }
  }
  // ONE launch when the jobs fit the kernel's parameter table (the 256-wide networks: 10 jobs), else several
  for (size_t first = 0; first < all.size(); first += tc::kDwMaxJobs) {
    const int nj = (int)std::min<size_t>(tc::kDwMaxJobs, all.size() - first);
    tc::DwJobs jobs{};
    jobs.n = nj;
    for (int i = 0; i < nj; ++i) jobs.j[i] = all[first + i];
    // CTAs per job proportional to the bytes it streams (largest-remainder rounding so that every SM gets a CTA)
    int wsum = 0;
    for (int i = 0; i < nj; ++i) wsum += weight[first + i];
    int cnt[tc::kDwMaxJobs], rem[tc::kDwMaxJobs], used = 0;
    for (int i = 0; i < nj; ++i) {
      const long long x = (long long)S->num_sms * weight[first + i];
      cnt[i] = std::max(1, (int)(x / wsum));
      rem[i] = (int)(x % wsum);
      used += cnt[i];
    }
    while (used < S->num_sms) {
      int best = 0;
      for (int i = 1; i < nj; ++i) if (rem[i] > rem[best]) best = i;
      cnt[best]++; rem[best] = -1; used++;
    }
    int begin = 0;
    for (int i = 0; i < nj; ++i) {
      const int per = std::max((int)round_up((rows + cnt[i] - 1) / cnt[i], 64), 64);
      jobs.j[i].rows_per_cta = per;
      jobs.j[i].cta_begin = begin;
      jobs.j[i].cta_count = (rows + per - 1) / per;          // (<= cnt[i]; trailing CTAs of the range would have no rows)
      begin += jobs.j[i].cta_count;
    }
    const int smem = tc::kDwStages * 8 * tc::kDwSlab + 256 + 1024;
    ProfScope prof(h, st, MARF_PROF_DW);
    launch_k(tc::k_tc_dw, begin, tc::kDwThreads, smem, st, jobs);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

// schedule of the chain kernels on CTA pairs: 1 = tile-staggered N = 256 (tc_chain.cuh: chain_role_staggered), 0 = two halves in lockstep
static int chain_sched() {
  static const int v = getenv("MARF_CHAIN_SCHED") ? atoi(getenv("MARF_CHAIN_SCHED")) : 0;
  return v;
}

// ---- fused chains (tc_chain.cuh): all hidden layers (+ output layer) of up to two MLPs in ONE launch
static int launch_chain(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows, bool forward) {
  Bf16State* S = h->bf16;
  ProfScope prof(h, st, forward ? MARF_PROF_CHAIN_FWD : MARF_PROF_CHAIN_DX);
  // CTA pairs (cluster of 2, cta_group::2 MMAs) by default; MARF_CHAIN_CL=1 selects the single-CTA variant (tests, A/B runs)
  const char* cl_env = getenv("MARF_CHAIN_CL");
  const int cl = (cl_env && atoi(cl_env) == 1) || S->num_sms < 2 ? 1 : 2;
  const bool stag = cl == 2 && chain_sched() == 1;          // (weight chunks: this CTA's 128 of 256 rows -> the 128-row boxes)
  tc::ChainJobs jobs{};
  jobs.n = n_chains;
  jobs.n_tiles = rows / 128;
  jobs.dbg = getenv("MARF_CHAIN_DBG") ? atoi(getenv("MARF_CHAIN_DBG")) : 0;
  for (int ci = 0; ci < n_chains; ++ci) {
    BfChain& B = *chains[ci];
    tc::ChainJob& J = jobs.c[ci];
    const int n = B.n;                       // n - 1 hidden layers + the thin output layer
    J.bits_ld = 256 / 32;
    if (forward) {
      J.tmIn = B.tmAct128[0];
      J.tmWout = cl == 2 ? B.tmWout8 : B.tmWout16;
      for (int u = 0; u < tc::kChUnits; ++u) {
        J.u[u].tmW = (cl == 2 && !stag) ? B.L[u].tmWk64 : B.L[u].tmWk128;
        J.u[u].tmOut = B.tmAct128[u + 1];
        J.u[u].bias = (u == 0 && B.col_off0 > 0) ? B.zero_bias : B.f32->bp[u];   // (class-table mode: b0 lives in Wk)
        J.u[u].bits = B.bits[u + 1];
      }
      J.bias_out = B.f32->bp[n - 1];
      J.logits = B.logits;
      J.k_out = B.L[n - 1].k_out;
    } else {
      // unit 0: dY[n-2] = (dlogits W_last) * mask(act[n-1]); unit u >= 1: layer l = n-1-u, dY[l-1] = (dY[l] W_l) * mask(act[l])
      J.tmIn = B.tmDL128;
      J.tmWout = B.tmWout16;                 // unused
      for (int u = 0; u < tc::kChUnits; ++u) {
        const int l = n - 1 - u;
        J.u[u].tmW = u == 0 ? ((cl == 2 && !stag) ? B.tmWlast64 : B.tmWlast128) : ((cl == 2 && !stag) ? B.L[l].tmWt64 : B.L[l].tmWt128);
        J.u[u].tmOut = B.tmDY128[l - 1];
        J.u[u].bias = nullptr;
        J.u[u].bits = B.bits[l];
      }
    }
  }
  const int group = 2 * cl;
  const int n_items = (jobs.n_tiles + group - 1) / group * n_chains;
  const int grid = cl * std::min(n_items, S->num_sms / cl);
  const int smem = tc::kChSmem + 1024;
  // diagnostics (MARF_CHAIN_TRACE=<n>): clock64 stamps of CTA 0 during the n-th chain launch
  static int calls = 0;
  const char* te = getenv("MARF_CHAIN_TRACE");
  const bool tracing = te && ++calls == atoi(te);
  if (tracing) {
    cudaMalloc(&jobs.trace, 2 * 4 * 2 * 16 * sizeof(long long));
    cudaMemset(jobs.trace, 0, 2 * 4 * 2 * 16 * sizeof(long long));
  }
  if (stag) {
    if (forward) launch_k_cluster(tc::k_tc_chain<tc::CH_FWD, 2, 1>, grid, tc::kChThreads, smem, st, 2, jobs);
    else launch_k_cluster(tc::k_tc_chain<tc::CH_DX, 2, 1>, grid, tc::kChThreads, smem, st, 2, jobs);
  } else if (cl == 2) {
    if (forward) launch_k_cluster(tc::k_tc_chain<tc::CH_FWD, 2>, grid, tc::kChThreads, smem, st, 2, jobs);
    else launch_k_cluster(tc::k_tc_chain<tc::CH_DX, 2>, grid, tc::kChThreads, smem, st, 2, jobs);
  } else {
    if (forward) launch_k(tc::k_tc_chain<tc::CH_FWD, 1>, grid, tc::kChThreads, smem, st, jobs);
    else launch_k(tc::k_tc_chain<tc::CH_DX, 1>, grid, tc::kChThreads, smem, st, jobs);
  }
  BF_LAUNCH(h);
  if (tracing) {
    cudaStreamSynchronize(st);
    long long tr[2 * 4 * 2 * 16];
    cudaMemcpy(tr, jobs.trace, sizeof(tr), cudaMemcpyDeviceToHost);
    cudaFree(jobs.trace);
    const long long t0 = tr[0];
    fprintf(stderr, "chain trace (%s): item unit half | mma: begin commit mid | g0: acc_full free ld0 ld1 sts arrive+store | g1: ...\n", forward ? "fwd" : "dx");
    for (int i = 0; i < 2 * 4 * 2; ++i) {
      fprintf(stderr, "%d %d %d |", i / 8, (i / 2) % 4, i % 2);
      for (int e = 0; e < 16; ++e) fprintf(stderr, " %6lld", tr[i * 16 + e] ? tr[i * 16 + e] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
  return MARF_OK;
}


// ---- the whole backward pass of the fused 256-wide networks in ONE launch (tc_bwd.cuh): the dX chains on the first CTA pairs,
// every dW / db GEMM on the others, dY handed over tile by tile through L2
static bool bwd_fused_enabled() {
  static const bool on = getenv("MARF_NO_BWD_FUSE") == nullptr;
  return on;
}
static int launch_bwd(marf_handle* h, cudaStream_t st, BfChain** chains, int n_chains, int rows) {
  Bf16State* S = h->bf16;
  ProfScope prof(h, st, MARF_PROF_BWD);
  tc::BwdJobs jobs{};
  const int n_tiles = rows / 128;
  const int total_pairs = S->num_sms / 2;
  jobs.rows = rows;
  jobs.chain.n = n_chains;
  jobs.chain.n_tiles = n_tiles;
  jobs.chain.ready = S->ready;
  jobs.chain.epoch = ++S->epoch;
  jobs.chain.interleave = 1;
  // dY tiles are stored with evict_last priority (they are consumed out of L2 by this launch) and dropped from L2 without a
  // write-back once consumed (discard.global.L2): -2.5 % launch time; MARF_BWD_STORE_LAST=0 / MARF_BWD_DISCARD=0 switch them off
  jobs.chain.store_last = getenv("MARF_BWD_STORE_LAST") ? atoi(getenv("MARF_BWD_STORE_LAST")) : 1;
  const bool discard = getenv("MARF_BWD_DISCARD") ? atoi(getenv("MARF_BWD_DISCARD")) != 0 : true;
  // (measured: no gain — 471 vs 463 us on config 2, 1325 vs 1259 us on config 4; the stage time does not follow the TMA bytes)
  const bool ldgsts = getenv("MARF_BWD_LDGSTS") ? atoi(getenv("MARF_BWD_LDGSTS")) != 0 : false;
  for (int ci = 0; ci < n_chains; ++ci) {
    BfChain& B = *chains[ci];
    tc::ChainJob& J = jobs.chain.c[ci];
    const int n = B.n;
    J.bits_ld = 256 / 32;
    // unit 0: dY[n-2] = (dlogits W_last) * mask(act[n-1]); unit u >= 1: layer l = n-1-u, dY[l-1] = (dY[l] W_l) * mask(act[l])
    J.tmIn = B.tmDL128;
    J.tmWout = B.tmWout16;                   // unused
    for (int u = 0; u < tc::kChUnits; ++u) {
      const int l = n - 1 - u;
      J.u[u].tmW = chain_sched() == 1 ? (u == 0 ? B.tmWlast128 : B.L[l].tmWt128) : (u == 0 ? B.tmWlast64 : B.L[l].tmWt64);
      J.u[u].tmOut = B.tmDY128[l - 1];
      J.u[u].bias = nullptr;
      J.u[u].bits = B.bits[l];
    }
  }
  // dW jobs: a 64-row stage costs a dW pair about the same whatever the job (four MMAs per issuer thread at ~90 cycles of issue
  // each against 128 / 64 cycles of tensor work for N = 256 / 128): equal weights, slightly less for the 64-wide layer-0 inputs
  double weight[tc::kBwdMaxJobs];
  auto add = [&](const CUtensorMap& tmDY, const CUtensorMap& tmX, const uint32_t* ready, int n_cols, int m_valid, int n_valid,
                 int ld_w, int do_bias, float* dW, float* db, int dy_cols, int x_cols) {
    tc::BwdDwJob& D = jobs.dw[jobs.n_dw];
    D.tmDY = tmDY; D.tmX = tmX; D.ready = ready; D.n_cols = n_cols; D.m_valid = m_valid; D.n_valid = n_valid;
    D.ld_w = ld_w; D.do_bias = do_bias; D.dW = dW; D.db = db; D.dy_cols = dy_cols; D.x_cols = x_cols;
    // a stage costs a pair the TMA boxes of its busier CTA (4 for the 256 x 256 jobs, 3 for the output layers and layer 0)
    int worst = 0;
    for (int r = 0; r < 2; ++r) {
      int nb = 0;
      for (int sl = 0; sl < 2; ++sl) nb += (r * 128 + sl * 64 < dy_cols) ? 1 : 0;
      for (int b = 0; b < n_cols / 128; ++b) nb += ((r * (n_cols / 128) + b) * 64 < x_cols) ? 1 : 0;
      worst = std::max(worst, nb);
    }
    weight[jobs.n_dw++] = worst / 4.0;
  };
  for (int ci = 0; ci < n_chains; ++ci) {
    BfChain& B = *chains[ci];
    Chain& F = *B.f32;
    const int n = B.n;
    if (jobs.n_dw + n > tc::kBwdMaxJobs) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 backward: too many layers");
    add(B.tmDL128, B.tmAct128[n - 1], nullptr, 256, B.L[n - 1].k_out, B.L[n - 1].k_in, F.ld_in[n - 1], 1, F.gWp[n - 1], F.gbp[n - 1], 8, 256);
    if (ldgsts) {
      jobs.dw[jobs.n_dw - 1].x_base = reinterpret_cast<const unsigned char*>(B.act[n - 1]);
      jobs.dw[jobs.n_dw - 1].x_pitch = B.ld[n - 1] * 2;
    }
    for (int l = n - 2; l >= 0; --l) {
      // dY[l] is the output of chain unit n-2-l
      const uint32_t* ready = S->ready + (size_t)(ci * tc::kChUnits + (n - 2 - l)) * n_tiles;
      if (l == 0 && B.col_off0 > 0)
        // class-table mode: the whole [256, 64] tile (uv columns and per-class sums) goes to the scratch that the tail launch
        // (k_unpack_table modes 1 / 2) turns into dW0 / db0
        add(B.tmDY128[0], B.tmAct128[0], ready, 128, B.L[0].k_out, 64, 64, 0, B.dW0x, F.gbp[0], 256, 64);
      else
        add(B.tmDY128[l], B.tmAct128[l], ready, l == 0 ? 128 : 256, B.L[l].k_out, B.L[l].k_in, F.ld_in[l], 1, F.gWp[l], F.gbp[l], 256,
            l == 0 ? 64 : 256);
      // MARF_BWD_LDGSTS=1 (experiment): 256-column X loaded by cp.async beside the TMA loads of dY
      if (ldgsts && l >= 1) {
        jobs.dw[jobs.n_dw - 1].x_base = reinterpret_cast<const unsigned char*>(B.act[l]);
        jobs.dw[jobs.n_dw - 1].x_pitch = B.ld[l] * 2;
      }
      // (dY[0] of a chain whose input gradient is needed is read again by the warp-gradient GEMM: kept)
      if (discard && !(l == 0 && B.need_dx0)) {
        jobs.dw[jobs.n_dw - 1].dy_base = reinterpret_cast<unsigned char*>(B.dY[l]);
        jobs.dw[jobs.n_dw - 1].dy_pitch = B.L[l].np * 2;
      }
    }
  }
  // chain pairs : dW pairs.  Measured on the B200 (profiles/r02_bwd_experiments.md): both roles run ~25 % slower side by side than
  // alone (they share L2 / HBM), a dW pair is bound by its operand loads (~0.85 us per 128-row stage whatever the job), and the
  // launch is fastest when the two roles finish together: 38 : 36 with two networks (10 jobs), 39 : 35 with one (5 jobs).
  // MARF_BWD_CHAIN_CLUSTERS overrides.
  double wsum = 0;
  for (int i = 0; i < jobs.n_dw; ++i) wsum += weight[i];
  int n_chain = (int)(total_pairs * (n_chains == 2 ? 0.515 : 0.53) + 0.5);
  if (const char* e = getenv("MARF_BWD_CHAIN_CLUSTERS")) n_chain = atoi(e);
  const int n_items = (n_tiles + 3) / 4 * n_chains;
  n_chain = std::max(1, std::min(n_chain, std::min(n_items, total_pairs - jobs.n_dw)));
  if (n_chains == 2 && n_chain > 1) n_chain &= ~1;            // an even count keeps every chain pair on one network
  const int n_dw_pairs = total_pairs - n_chain;
  if (n_dw_pairs < jobs.n_dw) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 backward: fewer CTA pairs than dW jobs");
  // pairs per job: proportional to the weights, largest-remainder rounding, at least one each
  int cnt[tc::kBwdMaxJobs], used = 0;
  double rem[tc::kBwdMaxJobs];
  for (int i = 0; i < jobs.n_dw; ++i) {
    const double x = n_dw_pairs * weight[i] / wsum;
    cnt[i] = std::max(1, (int)x);
    rem[i] = x - cnt[i];
    used += cnt[i];
  }
  while (used < n_dw_pairs) {
    int best = 0;
    for (int i = 1; i < jobs.n_dw; ++i) if (rem[i] > rem[best]) best = i;
    cnt[best]++; rem[best] -= 1.0; used++;
  }
  while (used > n_dw_pairs) {                                 // (the minimum of one pair per job overshot)
    int best = -1;
    for (int i = 0; i < jobs.n_dw; ++i) if (cnt[i] > 1 && (best < 0 || rem[i] < rem[best])) best = i;
    if (best < 0) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 backward: cannot place the dW jobs");
    cnt[best]--; rem[best] += 1.0; used--;
  }
  int begin = 0;
  for (int i = 0; i < jobs.n_dw; ++i) { jobs.dw[i].pair_begin = begin; jobs.dw[i].pair_count = cnt[i]; begin += cnt[i]; }
  jobs.n_chain_clusters = n_chain;
  jobs.prefetch_ahead = getenv("MARF_BWD_PREFETCH") ? atoi(getenv("MARF_BWD_PREFETCH")) : 0;
  const int grid = 2 * (n_chain + n_dw_pairs);
  // timing experiments (results invalid): MARF_BWD_DBG bit 1 = the dW pairs do not wait for the chain pairs, bit 2 = the chain
  // pairs do not publish their tiles
  const int dbg = getenv("MARF_BWD_DBG") ? atoi(getenv("MARF_BWD_DBG")) : 0;
  if (dbg & 1) for (int i = 0; i < jobs.n_dw; ++i) jobs.dw[i].ready = nullptr;
  if (dbg & 2) jobs.chain.ready = nullptr;
  // diagnostics (MARF_BWD_TRACE=<n>): begin / end time of every CTA pair of the n-th launch
  static int calls = 0;
  const char* te = getenv("MARF_BWD_TRACE");
  const bool tracing = te && ++calls == atoi(te);
  if (tracing) {
    cudaMalloc(&jobs.trace, (size_t)grid * sizeof(unsigned long long));
    cudaMemset(jobs.trace, 0, (size_t)grid * sizeof(unsigned long long));
  }
  if (chain_sched() == 1) launch_k_cluster(tc::k_tc_bwd<1>, grid, tc::kChThreads, tc::kChSmem + 1024, st, 2, jobs);
  else launch_k_cluster(tc::k_tc_bwd<0>, grid, tc::kChThreads, tc::kChSmem + 1024, st, 2, jobs);
  BF_LAUNCH(h);
  if (tracing) {
    cudaStreamSynchronize(st);
    std::vector<unsigned long long> tr(grid);
    cudaMemcpy(tr.data(), jobs.trace, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(jobs.trace);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid / 2; ++c) t0 = std::min(t0, tr[2 * c]);
    double c_end = 0, d_end = 0;
    for (int c = 0; c < grid / 2; ++c) (c < n_chain ? c_end : d_end) = std::max(c < n_chain ? c_end : d_end, (tr[2 * c + 1] - t0) * 1e-3);
    fprintf(stderr, "k_tc_bwd trace: %d chain pairs (last ends %.1f us), %d dW pairs (last ends %.1f us), rows %d; per pair begin/end in us\n",
            n_chain, c_end, n_dw_pairs, d_end, rows);
    for (int c = 0; c < grid / 2; ++c) {
      int job = -1;
      if (c >= n_chain) for (int i = 0; i < jobs.n_dw; ++i) if (c - n_chain >= jobs.dw[i].pair_begin) job = i;
      fprintf(stderr, "  pair %2d %s job %2d  %8.1f %8.1f\n", c, c < n_chain ? "chain" : "dW   ", job, (tr[2 * c] - t0) * 1e-3, (tr[2 * c + 1] - t0) * 1e-3);
    }
  }
  return MARF_OK;
}

static int thin_fwd(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* W, const float* bias) {
  int l = B.n - 1;
  int width = B.L[l].k_in, out = B.L[l].k_out;
  int blocks = std::min((rows + 63) / 64, h->bf16->num_sms * 8);
  if (out == 3 && width == 256) launch_k(k_thin_fwd<3, 1>, blocks, 256, 0, st, rows, B.act[l], B.ld[l], W, bias, B.logits);
  else if (out == 3 && width == 512) launch_k(k_thin_fwd<3, 2>, blocks, 256, 0, st, rows, B.act[l], B.ld[l], W, bias, B.logits);
  else if (out == 1 && width == 256) launch_k(k_thin_fwd<1, 1>, blocks, 256, 0, st, rows, B.act[l], B.ld[l], W, bias, B.logits);
  else if (out == 1 && width == 512) launch_k(k_thin_fwd<1, 2>, blocks, 256, 0, st, rows, B.act[l], B.ld[l], W, bias, B.logits);
  else return fail(h, MARF_ERR_UNSUPPORTED, "bf16: output layer must be 3- or 1-wide over 256/512 features");
  BF_LAUNCH(h);
  return MARF_OK;
}

// last layer backward: dW/db (fp32 atomics into the padded fp32 twin) and dY of the previous layer
static int thin_bwd(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* W) {
  Bf16State* S = h->bf16;
  Chain& F = *B.f32;
  int l = B.n - 1;
  int width = B.L[l].k_in, out = B.L[l].k_out;
  int rpb = std::max(256, (rows + 2 * S->num_sms - 1) / (2 * S->num_sms));
  dim3 blk(width / 8, 8);
  int nblk = (rows + rpb - 1) / rpb;
  int smem = 8 * (out * width + out) * sizeof(float);
  int per_row = width / 8;
  int dxthreads = 256 / per_row * per_row;             // threads per block: multiple of per_row
  int dxblocks = std::min((rows * per_row + dxthreads - 1) / dxthreads, S->num_sms * 16);
  const bool tc_dw = width == 256;           // then launch_dw_all carries the output layer's dW/db as a tensor-core job
  if (B.Wlast_t) {
    // dY[n-2] = (dlogits W_last) * relu_mask: the same GEMM kernel as the hidden dX layers, K = 64 (3 or 1 real columns)
    tc::GemmJobs one{};
    one.n = 1;
    tc::GemmJob& J = one.j[0];
    J.tmA = B.tmDL128;
    J.tmW = B.tmWlast_t;
    J.tmOut = B.tmDY128[l - 1];
    J.p.n_tiles = rows / 128;
    J.p.k_chunks = 1;
    J.p.bits_in = B.bits[l];
    J.p.bits_ld = B.ld[l] / 32;
    J.p.reverse = 0;
    J.p.load_policy = tc::kEvictNormal;        // dlogits are read again by the dW pass
    J.p.store_policy = tc::kEvictLast;
    J.cta_begin = 0;
    J.cta_count = std::min(J.p.n_tiles, S->num_sms);
    int smem_g = tc::gemm_smem(256, 1, true).total + 1024;
    launch_k(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, J.cta_count, tc::kThreads, smem_g, st, one);
    BF_LAUNCH(h);
    return MARF_OK;
  }
  if (out == 3) {
    if (!tc_dw) {
      launch_k(k_thin_dw<3>, nblk, blk, smem, st, rows, width, B.dlogits, B.act[l], B.ld[l], F.gWp[l], F.ld_in[l], F.gbp[l], rpb);
      BF_LAUNCH(h);
    }
    launch_k(k_thin_dx<3>, dxblocks, dxthreads, 0, st, rows, width, B.dlogits, W, B.bits[l], B.ld[l] / 32, B.dY[l - 1], B.L[l - 1].np);
  } else {
    if (!tc_dw) {
      launch_k(k_thin_dw<1>, nblk, blk, smem, st, rows, width, B.dlogits, B.act[l], B.ld[l], F.gWp[l], F.ld_in[l], F.gbp[l], rpb);
      BF_LAUNCH(h);
    }
    launch_k(k_thin_dx<1>, dxblocks, dxthreads, 0, st, rows, width, B.dlogits, W, B.bits[l], B.ld[l] / 32, B.dY[l - 1], B.L[l - 1].np);
  }
  BF_LAUNCH(h);
  return MARF_OK;
}

// one launch: bias (fp32) + bf16 forward / transposed weights of every tensor-core layer
static size_t chain_twin_floats(const Chain& C) {       // the padded fp32 gradient twins of a chain are one allocation
  size_t t = 0;
  for (int l = 0; l < C.n; ++l) t += (size_t)C.ld_out[l] * C.ld_in[l] + C.ld_out[l];
  return t;
}

// fused: the launch also computes the homographies and the mask head's class table and zeroes the step's accumulators
// (single-call marf_step); otherwise those stay separate launches / memsets (two-phase entry points, multi-chunk sweeps)
static int pack_all(marf_handle* h, cudaStream_t st, const marf_step_io* io, bool fused = false) {
  Bf16State* S = h->bf16;
  PackTable t{};
  t.n = 0;
  int max_tot = 0;
  auto add = [&](const float* src, void* dst, int rows, int cols, int prow, int pcol, int mode, int src_ld = -1, int col_off = 0,
                 int wcols = 0) {
    PackEntry& e = t.e[t.n++];
    e.src = src; e.dst = dst; e.rows = rows; e.cols = cols; e.prow = prow; e.pcol = pcol; e.mode = mode;
    e.src_ld = src_ld < 0 ? cols : src_ld; e.col_off = col_off; e.wcols = wcols;
    max_tot = std::max(max_tot, prow * pcol);
  };
  BfChain* chains[2] = {&S->img, h->cfg.mask_mode == MARF_MASK_IMPLICIT ? &S->msk : nullptr};
  const float* const* Ws[2] = {io->mlp_w, io->mask_w};
  const float* const* bs[2] = {io->mlp_b, io->mask_b};
  for (int ci = 0; ci < 2; ++ci) {
    if (!chains[ci]) continue;
    BfChain& B = *chains[ci];
    Chain& F = *B.f32;
    for (int l = 0; l < B.n; ++l) {
      if (!Ws[ci][l] || !bs[ci][l]) return fail(h, MARF_ERR_INVALID, "null layer parameter");
      if (t.n + 4 > kMaxPack) return fail(h, MARF_ERR_UNSUPPORTED, "too many layers");
      add(bs[ci][l], F.bp[l], 1, F.k_out[l], 1, F.ld_out[l], 0);
      if (B.L[l].thin && B.Wlast_t) add(Ws[ci][l], B.Wlast_t, B.L[l].k_out, B.L[l].k_in, B.L[l].k_in, 64, 2);
      if (B.L[l].thin && B.Wout16) add(Ws[ci][l], B.Wout16, B.L[l].k_out, B.L[l].k_in, 16, B.L[l].k_in, 3);
      if (B.L[l].thin) continue;
      // (layer 0 of the mask head only multiplies the uv columns: a window of the caller's [k_out, k_in] matrix)
      const int off = l == 0 ? B.col_off0 : 0;
      // (class-table mode: columns >= k_in of Wk belong to the table k_mask_table / the fused prologue writes: never touched here)
      add(Ws[ci][l], B.L[l].Wk, B.L[l].k_out, B.L[l].k_in, B.L[l].np, B.L[l].kp, 1, F.k_in[l], off, off > 0 ? B.L[l].k_in : 0);
      add(Ws[ci][l], B.L[l].Wt, B.L[l].k_out, B.L[l].k_in, B.L[l].kp, B.L[l].np, 2, F.k_in[l], off);
    }
  }
  const bool implicit = h->cfg.mask_mode == MARF_MASK_IMPLICIT;
  t.fused = fused ? 1 : 0;
  if (fused) {
    t.warp = io->warp; t.n_patches = h->cfg.batch_global; t.Hm = h->Hm;
    if (implicit) {
      t.mW0 = io->mask_w[0]; t.mb0 = io->mask_b[0]; t.embed = io->embed; t.mk_in = h->msk.k_in[0]; t.mk_out = h->msk.k_out[0];
      t.edim = h->cfg.mask_embed_dim; t.k_uv = S->msk.L[0].k_in; t.mWk = S->msk.L[0].Wk; t.mldk = S->msk.L[0].kp;
    }
    int z = 0;
    t.zptr[z] = h->img.gWp[0]; t.zbytes[z++] = chain_twin_floats(h->img) * sizeof(float);
    if (implicit) { t.zptr[z] = h->msk.gWp[0]; t.zbytes[z++] = chain_twin_floats(h->msk) * sizeof(float); }
    t.zptr[z] = h->G; t.zbytes[z++] = (size_t)h->cfg.batch * 9 * sizeof(double);
    if (implicit) { t.zptr[z] = S->msk.dW0x; t.zbytes[z++] = 256 * 64 * sizeof(float); }
    t.zptr[z] = io->loss_sums; t.zbytes[z++] = MARF_N_SUMS * sizeof(double);
  }
  dim3 grid(std::min((max_tot + 255) / 256, 64), t.n + (fused ? 3 : 0));
  launch_k(k_pack_table, grid, 256, 0, st, t);
  BF_LAUNCH(h);
  return MARF_OK;
}

static int unpack_all(marf_handle* h, cudaStream_t st, const marf_step_io* io) {
  Bf16State* S = h->bf16;
  UnpackTable t{};
  t.n = 0;
  if (!io->g_warp) return fail(h, MARF_ERR_INVALID, "missing g_warp");
  t.warp = io->warp; t.G = h->G; t.g_warp = io->g_warp;
  t.patch_offset = h->cfg.patch_offset; t.n_local = h->cfg.batch; t.n_global = h->cfg.batch_global;
  int max_tot = 0;
  BfChain* chains[2] = {&S->img, h->cfg.mask_mode == MARF_MASK_IMPLICIT ? &S->msk : nullptr};
  float* const* gW[2] = {io->g_mlp_w, io->g_mask_w};
  float* const* gb[2] = {io->g_mlp_b, io->g_mask_b};
  for (int ci = 0; ci < 2; ++ci) {
    if (!chains[ci]) continue;
    if (!gW[ci] || !gb[ci]) return fail(h, MARF_ERR_INVALID, "missing gradient pointers");
    Chain& F = *chains[ci]->f32;
    for (int l = 0; l < F.n; ++l) {
      UnpackEntry& a = t.e[t.n++];
      a.src = F.gWp[l]; a.dst = gW[ci][l]; a.rows = F.k_out[l]; a.cols = F.k_in[l]; a.pcol = F.ld_in[l];
      UnpackEntry& b = t.e[t.n++];
      b.src = F.gbp[l]; b.dst = gb[ci][l]; b.rows = 1; b.cols = F.k_out[l]; b.pcol = F.ld_out[l];
      if (l == 0 && chains[ci]->col_off0 > 0) {
        // mask head layer 0 (class-table mode): from the [k_out, 64] scratch the dW launch accumulated
        a.src = chains[ci]->dW0x; a.mode = 1; a.embed = io->embed; a.edim = h->cfg.mask_embed_dim; a.k_uv = chains[ci]->L[0].k_in;
        b.src = chains[ci]->dW0x; b.mode = 2; b.embed = io->embed; b.edim = h->cfg.mask_embed_dim; b.k_uv = chains[ci]->L[0].k_in;
      }
      max_tot = std::max(max_tot, F.k_out[l] * F.k_in[l]);
    }
  }
  dim3 grid(std::min((max_tot + 255) / 256, 64), t.n + 1);
  launch_k(k_unpack_table, grid, 256, 0, st, t);
  BF_LAUNCH(h);
  return MARF_OK;
}

// ------------------------------------------------------------------------------------------------ the step
static PxRange bf_chunk(const marf_handle* h, int ci) {
  PxRange rg;
  rg.first = (long long)ci * h->chunk;
  rg.count = (int)std::min<long long>(h->chunk, h->n_local - rg.first);
  rg.padded = (int)round_up(rg.count, 128);
  return rg;
}

static int bf_forward_chunk(marf_handle* h, const marf_step_io* io, cudaStream_t st, int ci, bool stats) {
  Bf16State* S = h->bf16;
  const marf_config& c = h->cfg;
  PxRange rg = bf_chunk(h, ci);
  bool implicit = c.mask_mode == MARF_MASK_IMPLICIT;
  {
    dim3 eg((rg.padded + 127) / 128);
    if (h->geo.L == 8) launch_k(k_encode_bf16<8>, eg, 128, 0, st, h->geo, rg, h->Hm, S->img.act[0], S->img.ld[0]);
    else if (h->geo.L == 10) launch_k(k_encode_bf16<10>, eg, 128, 0, st, h->geo, rg, h->Hm, S->img.act[0], S->img.ld[0]);
    else if (h->geo.L == 4) launch_k(k_encode_bf16<4>, eg, 128, 0, st, h->geo, rg, h->Hm, S->img.act[0], S->img.ld[0]);
    else launch_k(k_encode_bf16<0>, eg, 128, 0, st, h->geo, rg, h->Hm, S->img.act[0], S->img.ld[0]);
  }
  BF_LAUNCH(h);
  if (implicit) {
    if (!(h->feats_valid && h->n_chunks == 1)) {
      // the class table needs trunc(rgb) in {0,1} (true for [0,1] images); every chunk is checked on device and the count
      // of offending indices reaches the host with the loss sums (MARF_BAD_INDEX) — no synchronisation inside the step
      launch_k(k_mask_uv_cls, (rg.padded + 127) / 128, 128, 0, st, h->geo, rg, io->rgb, c.mask_uv_freqs, S->msk.act[0], h->bad_index);
      BF_LAUNCH(h);
      h->feats_valid = h->n_chunks == 1;
    }
    if (!S->table_done) {                         // (the fused prologue of a single-call step has already written it)
      launch_k(k_mask_table, (8 * h->msk.k_out[0] * 32 + 255) / 256, 256, 0, st, io->mask_w[0], io->mask_b[0], io->embed,
               h->msk.k_in[0], h->msk.k_out[0], c.mask_embed_dim, S->msk.L[0].k_in, S->msk.L[0].Wk, S->msk.L[0].kp);
      BF_LAUNCH(h);
    }
  }
  // chain after chain, each followed at once by its output layer: the last hidden activation is still in L2
  BfChain* c_img[1] = {&S->img};
  BfChain* c_msk[1] = {&S->msk};
  int rc = MARF_OK;
  if (S->img.fused && (!implicit || S->msk.fused)) {
    // layer-fused: every hidden layer and the output layer of both networks in one launch, activations resident in SMEM
    BfChain* both[2] = {&S->img, &S->msk};
    rc = launch_chain(h, st, both, implicit ? 2 : 1, rg.padded, true);
    if (rc) return rc;
  } else {
    rc = launch_forward_all(h, st, c_img, 1, rg.padded);
    if (rc) return rc;
    rc = thin_fwd(h, st, S->img, rg.padded, io->mlp_w[c.n_layers - 1], io->mlp_b[c.n_layers - 1]);
    if (rc) return rc;
    if (implicit) {
      rc = launch_forward_all(h, st, c_msk, 1, rg.padded);
      if (rc) return rc;
      rc = thin_fwd(h, st, S->msk, rg.padded, io->mask_w[c.mask_n_layers - 1], io->mask_b[c.mask_n_layers - 1]);
      if (rc) return rc;
    }
  }
  if (stats) {
    LossArgs a;
    a.mask_mode = c.mask_mode;
    a.logits = S->img.logits; a.ld = 4;
    a.mlogits = implicit ? S->msk.logits : nullptr; a.mld = 4;
    a.rgb = io->rgb; a.masks = io->masks;
    a.rgb_pred = io->rgb_pred ? io->rgb_pred : h->pred_rgb;
    a.mask_pred = io->mask_pred ? io->mask_pred : h->pred_mask;
    a.bad_index = implicit ? h->bad_index : nullptr;
    launch_k(k_loss_stats, std::min((rg.padded + 255) / 256, 592), 256, 0, st, h->geo, rg, a, io->loss_sums);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

// coef == nullptr: k_loss_grad resolves the loss coefficients itself from the (all-reduced) sums of the forward pass
static int bf_backward_chunk(marf_handle* h, const marf_step_io* io, cudaStream_t st, int ci, const LossCoef* coef = nullptr) {
  Bf16State* S = h->bf16;
  const marf_config& c = h->cfg;
  PxRange rg = bf_chunk(h, ci);
  bool implicit = c.mask_mode == MARF_MASK_IMPLICIT;
  GradArgs ga;
  ga.l.mask_mode = c.mask_mode;
  ga.l.logits = S->img.logits; ga.l.ld = 4;
  ga.l.mlogits = implicit ? S->msk.logits : nullptr; ga.l.mld = 4;
  ga.l.rgb = io->rgb; ga.l.masks = io->masks; ga.l.rgb_pred = nullptr; ga.l.mask_pred = nullptr; ga.l.bad_index = nullptr;
  ga.c_rgb = io->c_rgb; ga.c_mask = io->c_mask; ga.c_edge = io->c_edge;
  ga.edge_pred = (implicit && c.use_edges) ? (io->edge_pred ? io->edge_pred : h->edge_pred) : nullptr;
  ga.edge_label = io->edges; ga.label_channels = c.edge_label_channels > 0 ? c.edge_label_channels : 1;
  ga.dlogits = S->img.dlogits; ga.dld = 4;
  ga.dmlogits = implicit ? S->msk.dlogits : nullptr; ga.dmld = 4;
  ga.dl_bf16 = S->img.dl16;
  ga.dml_bf16 = implicit ? S->msk.dl16 : nullptr;
  ga.sums = io->loss_sums; ga.norm_rgb = io->norm_rgb; ga.norm_edge = io->norm_edge; ga.use_edges = c.use_edges;
  launch_k(k_loss_grad, (rg.padded + 127) / 128, 128, 0, st, h->geo, rg, ga, coef);
  BF_LAUNCH(h);
  BfChain* c_img[1] = {&S->img};
  BfChain* c_msk[1] = {&S->msk};
  int rc = MARF_OK;
  if (S->img.fused && (!implicit || S->msk.fused)) {
    BfChain* both[2] = {&S->img, &S->msk};
    if (bwd_fused_enabled() && S->num_sms >= 2 * (2 + 2 * tc::kBwdMaxJobs)) {
      rc = launch_bwd(h, st, both, implicit ? 2 : 1, rg.padded);          // dX chains + every dW / db GEMM, one launch
      if (rc) return rc;
      return launch_dx0(h, st, c_img, 1, rg.padded, rg);                  // dX0 + encoding backward -> per-patch G
    }
    rc = launch_chain(h, st, both, implicit ? 2 : 1, rg.padded, false);   // dlogits -> dY[n-2] -> ... -> dY[0], one launch
    if (rc) return rc;
    rc = launch_dx0(h, st, c_img, 1, rg.padded, rg);                      // dX0 + encoding backward -> per-patch G
    if (rc) return rc;
  } else {
    rc = thin_bwd(h, st, S->img, rg.padded, io->mlp_w[c.n_layers - 1]);
    if (rc) return rc;
    rc = launch_dx_all(h, st, c_img, 1, rg.padded, rg);       // ends with dX0 + encoding backward -> per-patch G
    if (rc) return rc;
    if (implicit) {
      rc = thin_bwd(h, st, S->msk, rg.padded, io->mask_w[c.mask_n_layers - 1]);
      if (rc) return rc;
      rc = launch_dx_all(h, st, c_msk, 1, rg.padded, rg);
      if (rc) return rc;
    }
  }
  BfChain* chains[2] = {&S->img, &S->msk};
  return launch_dw_all(h, st, chains, implicit ? 2 : 1, rg.padded, rg.first);
}

// shared with api.cu
int engine_begin_step(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool pack_fp32, bool sl3 = true);
int engine_edge_pass(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int engine_begin_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int engine_finish_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool unpack);

// fused = the single-call step: one prologue launch packs the weights, computes H and the class table and zeroes every
// accumulator of the step (forward AND backward), so that no memset node sits between the launches
static int bf16_forward_impl(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool fused) {
  const marf_config& c = h->cfg;
  int rc = engine_begin_step(h, io, st, false, !fused);   // schedule, data caches (+ H matrices unless fused)
  if (rc) return rc;
  rc = pack_all(h, st, io, fused);
  if (rc) return rc;
  h->bf16->table_done = fused;
  h->bf16->accs_zeroed = fused;
  if (!fused) BF_TRY(h, cudaMemsetAsync(io->loss_sums, 0, MARF_N_SUMS * sizeof(double), st));
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    rc = bf_forward_chunk(h, io, st, ci, true);
    if (rc) return rc;
  }
  if (c.use_edges) {
    rc = engine_edge_pass(h, io, st);
    if (rc) return rc;
  }
  h->acts_valid = h->n_chunks == 1;
  h->bf16->table_done = false;
  return MARF_OK;
}
// (two-phase entry points: the forward's prologue launch also zeroes what the backward pass accumulates into; the backward
//  pass skips its memsets when it finds them done)
int bf16_forward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  return bf16_forward_impl(h, io, st, h->n_chunks == 1 && !getenv("MARF_NO_FUSED_PROLOGUE"));
}

static int bf16_backward_impl(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool fused) {
  const marf_config& c = h->cfg;
  int rc = MARF_OK;
  const bool implicit = c.mask_mode == MARF_MASK_IMPLICIT;
  if (!fused) {
    rc = engine_begin_backward(h, io, st);
    if (rc) return rc;
    if (implicit) BF_TRY(h, cudaMemsetAsync(h->bf16->msk.dW0x, 0, 256 * 64 * sizeof(float), st));
  }
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    if (!h->acts_valid) {
      rc = bf_forward_chunk(h, io, st, ci, false);
      if (rc) return rc;
    }
    rc = bf_backward_chunk(h, io, st, ci);
    if (rc) return rc;
  }
  h->acts_valid = false;
  // (mask head layer 0, the sl(3) adjoint and the hand-over of every gradient: one launch)
  return unpack_all(h, st, io);
}
int bf16_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  const bool zeroed = h->bf16->accs_zeroed;
  h->bf16->accs_zeroed = false;
  return bf16_backward_impl(h, io, st, zeroed);
}

// Several chunks and a normaliser that does not depend on the forward pass (no masks / disk masks: the mask head is the
// only consumer of forward-dependent coefficients): forward + backward chunk by chunk in ONE sweep.  The generic path
// (bf16_forward over all chunks for the statistics, then bf16_backward re-running the forward of every chunk because only one
// chunk's activations are resident) costs a second forward pass.
static int bf16_step_sweep(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  const marf_config& c = h->cfg;
  int rc = engine_begin_step(h, io, st, false);
  if (rc) return rc;
  rc = pack_all(h, st, io);
  if (rc) return rc;
  BF_TRY(h, cudaMemsetAsync(io->loss_sums, 0, MARF_N_SUMS * sizeof(double), st));
  rc = engine_begin_backward(h, io, st);
  if (rc) return rc;
  launch_k(k_loss_coef_static, 1, 1, 0, st, h->sums_static, c.mask_mode, io->norm_rgb, h->n_local, h->coef);
  BF_LAUNCH(h);
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    rc = bf_forward_chunk(h, io, st, ci, true);
    if (rc) return rc;
    rc = bf_backward_chunk(h, io, st, ci, h->coef);
    if (rc) return rc;
  }
  if (c.use_edges) {                                   // (loss value only: without the mask head the edge term has no gradient)
    rc = engine_edge_pass(h, io, st);
    if (rc) return rc;
  }
  h->acts_valid = false;
  return unpack_all(h, st, io);
}

int bf16_step(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  if (h->n_chunks > 1 && h->cfg.mask_mode != MARF_MASK_IMPLICIT && !getenv("MARF_NO_SWEEP")) return bf16_step_sweep(h, io, st);
  const bool fused = h->n_chunks == 1 && !getenv("MARF_NO_FUSED_PROLOGUE");
  int rc = bf16_forward_impl(h, io, st, fused);
  if (rc) return rc;
  h->bf16->accs_zeroed = false;
  return bf16_backward_impl(h, io, st, fused);
}

}  // namespace marf

// ------------------------------------------------------------------------------------------------ diagnostics
// Copy a resident bf16 buffer of the last chunk to fp32: which 0 = act[layer] (layer input), 1 = dY[layer].  tests/ only.
extern "C" int marf_debug_read_bf16(marf_handle* h, int chain, int which, int layer, float* out, long long rows, void* stream) {
  using namespace marf;
  if (!h || !h->bf16) return MARF_ERR_INVALID;
  BfChain& B = chain == 0 ? h->bf16->img : h->bf16->msk;
  if (layer < 0 || layer >= B.n) return MARF_ERR_INVALID;
  const bf16* src = which == 0 ? B.act[layer] : B.dY[layer];
  const int ld = which == 0 ? B.ld[layer] : B.L[layer].np;
  if (!src || rows > h->chunk) return MARF_ERR_INVALID;
  const long long tot = rows * ld;
  launch_k(k_bf16_to_f32, (unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream, tot, src, out);
  return cudaStreamSynchronize((cudaStream_t)stream) == cudaSuccess ? MARF_OK : MARF_ERR_CUDA;
}

// One tensor-core kernel on caller-provided fp32 device arrays (rounded to bf16 inside), fp32 result.
//   mode 0: out[rows,N] = relu(A[rows,K] W[N,K]^T + aux[N])          (k_tc_gemm, EPI_BIAS_RELU; bf16-rounded output)
//   mode 1: out[rows,N] = (A W^T) * (aux[rows,N] > 0)                 (k_tc_gemm, EPI_RELU_MASK; bf16-rounded output)
//   mode 2: out[rows,64] = A[rows,K] W[64,K]^T                        (k_tc_gemm, EPI_PLAIN_F32)
//   mode 3: out[N(out),K(in)] = A[rows,N]^T aux[rows,K], then out[N*K + o] = colsum(A)[o]   (k_tc_dw; A = dY, aux = X)
extern "C" int marf_tc_selftest(int device, int mode, int rows, int K, int N, const float* A, const float* W,
                                const float* aux, float* out, void* stream) {
  using namespace marf;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaSetDevice(device) != cudaSuccess) return MARF_ERR_CUDA;
  if (rows % 128 || K % 64 || N % 64) return MARF_ERR_INVALID;
  marf_handle tmp;                 // only err/launches/allocs are touched
  Bf16State S;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return MARF_ERR_CUDA;
  S.encode = (EncodeTiledFn)fn;
  const int big = 232448;
  cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_PLAIN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  auto to_bf = [&](const float* src, int r, int c) -> bf16* {
    bf16* d = nullptr;
    if (cudaMalloc(&d, (size_t)r * c * 2) != cudaSuccess) return nullptr;
    int tot = r * c;
    launch_k(k_pack_bf16, (tot + 255) / 256, 256, 0, st, src, r, c, d, r, c, 0);
    return d;
  };
  int rc = MARF_OK;
  bf16 *dA = nullptr, *dW = nullptr, *dAux = nullptr, *dOut = nullptr;
  float* dBias = nullptr;
  CUtensorMap tA, tW, tO, tM;
  if (mode <= 2) {
    dA = to_bf(A, rows, K);
    dW = to_bf(W, N, K);
    if (!dA || !dW) return MARF_ERR_CUDA;
    int n_tile = mode == 2 ? 64 : ((K > 256 && N > 128) ? 128 : std::min(N, 256));
    if (mode == 2 && N != 64) return MARF_ERR_INVALID;
    if (N % n_tile) return MARF_ERR_INVALID;
    rc = make_tmap(&tmp, &S, &tA, dA, rows, K, 128);
    if (!rc) rc = make_tmap(&tmp, &S, &tW, dW, N, K, n_tile);
    tc::GemmJobs jobs{};
    jobs.n = N / n_tile;
    tc::GemmParams p{};
    p.n_tiles = rows / 128;
    p.k_chunks = K / 64;
    p.reverse = (rows / 128) & 1;
    p.load_policy = tc::kEvictFirst;
    p.store_policy = tc::kEvictLast;
    const int per_job = std::min(p.n_tiles, std::max(1, 148 / jobs.n));
    const int grid = per_job * jobs.n;
    auto fill = [&](const CUtensorMap& a_, const CUtensorMap& w_, const CUtensorMap& o_) {
      for (int t = 0; t < jobs.n; ++t) {
        jobs.j[t].tmA = a_; jobs.j[t].tmW = w_; jobs.j[t].tmOut = o_; jobs.j[t].p = p; jobs.j[t].n0 = t * n_tile;
        jobs.j[t].cta_begin = t * per_job; jobs.j[t].cta_count = per_job;
      }
    };
    if (mode == 2) {
      p.out_f32 = out; p.ld_out = 64; p.n_store = 64;
      int smem = tc::gemm_smem(64, p.k_chunks, false).total + 1024;
      fill(tA, tW, tA);
      if (!rc) launch_k(tc::k_tc_gemm<64, tc::EPI_PLAIN_F32>, grid, tc::kThreads, smem, st, jobs);
    } else {
      if (cudaMalloc(&dOut, (size_t)rows * N * 2) != cudaSuccess) return MARF_ERR_CUDA;
      if (!rc) rc = make_tmap(&tmp, &S, &tO, dOut, rows, N, 128);
      int smem = tc::gemm_smem(n_tile, p.k_chunks, true).total + 1024;
      long long* dTrace = nullptr;
      const int n_iter_trace = (p.n_tiles + per_job - 1) / per_job;
      if (mode == 0 && getenv("MARF_TC_TRACE")) {
        cudaMalloc(&dTrace, (size_t)n_iter_trace * 16 * sizeof(long long));
        cudaMemset(dTrace, 0, (size_t)n_iter_trace * 16 * sizeof(long long));
        p.trace = dTrace;
      }
      if (mode == 0) {
        p.bias = aux;
        fill(tA, tW, tO);
        if (!rc) {
          if (n_tile == 256) launch_k(tc::k_tc_gemm<256, tc::EPI_BIAS_RELU>, grid, tc::kThreads, smem, st, jobs);
          else if (n_tile == 128) launch_k(tc::k_tc_gemm<128, tc::EPI_BIAS_RELU>, grid, tc::kThreads, smem, st, jobs);
          else rc = MARF_ERR_INVALID;
        }
      } else {
        uint32_t* dBits = nullptr;
        if (cudaMalloc(&dBits, (size_t)rows * (N / 32) * 4) != cudaSuccess) return MARF_ERR_CUDA;
        launch_k(k_make_bits, (rows * (N / 32) + 255) / 256, 256, 0, st, rows, N, aux, dBits);
        p.bits_in = dBits;
        p.bits_ld = N / 32;
        fill(tA, tW, tO);
        if (!rc) {
          if (n_tile == 256) launch_k(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, grid, tc::kThreads, smem, st, jobs);
          else if (n_tile == 128) launch_k(tc::k_tc_gemm<128, tc::EPI_RELU_MASK>, grid, tc::kThreads, smem, st, jobs);
          else rc = MARF_ERR_INVALID;
        }
        cudaStreamSynchronize(st);
        cudaFree(dBits);
      }
      if (!rc) {
        long long tot = (long long)rows * N;
        launch_k(k_bf16_to_f32, (unsigned)((tot + 255) / 256), 256, 0, st, tot, dOut, out);
      }
      if (dTrace) {
        cudaStreamSynchronize(st);
        std::vector<long long> tr((size_t)n_iter_trace * 16);
        cudaMemcpy(tr.data(), dTrace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        long long t00 = tr[1];
        fprintf(stderr, "trace (cycles rel. to first MMA start): iter | prod_issued mma_start mma_committed | g0: acc_full ld0 st0 ld1 st1 | g1: acc_full ld0 st0 ld1 st1\n");
        for (int i = 0; i < n_iter_trace; ++i) {
          fprintf(stderr, "%3d |", i);
          for (int k = 0; k < 15; ++k) fprintf(stderr, " %7lld", tr[i * 16 + k] ? tr[i * 16 + k] - t00 : -1);
          fprintf(stderr, "\n");
        }
        cudaFree(dTrace);
      }
    }
  } else if (mode == 3) {
    dA = to_bf(A, rows, N);       // dY [rows, out=N]
    dAux = to_bf(aux, rows, K);   // X  [rows, in=K]
    if (!dA || !dAux) return MARF_ERR_CUDA;
    rc = make_tmap(&tmp, &S, &tA, dA, rows, N, 64);
    if (!rc) rc = make_tmap(&tmp, &S, &tM, dAux, rows, K, 64);
    int n_tile = K >= 256 ? 256 : 64;
    if (n_tile == 64 && K != 64) return MARF_ERR_INVALID;
    if (N > 256) return MARF_ERR_INVALID;
    int n_tiles_n = (K + n_tile - 1) / n_tile;
    int ctas = std::max(1, 148 / n_tiles_n);
    int per = std::max((int)round_up((rows + ctas - 1) / ctas, 64), 64);
    ctas = (rows + per - 1) / per;
    // out holds [N, K] weights followed by [N] bias sums
    cudaMemsetAsync(out, 0, ((size_t)N * K + N) * sizeof(float), st);
    tc::DwJobs jobs{};
    jobs.n = n_tiles_n;
    for (int t = 0; t < n_tiles_n; ++t) {
      tc::DwJob& J = jobs.j[t];
      J.tmDY = tA; J.tmX = tM; J.rows = rows; J.rows_per_cta = per; J.cta_begin = t * ctas; J.cta_count = ctas; J.n_tile = n_tile;
      J.m_halves = (N + 127) / 128; J.m_valid = N;
      J.n_valid = K; J.n0 = t * n_tile; J.ld_w = K; J.do_bias = t == 0; J.dW = out; J.db = out + (size_t)N * K;
    }
    int smem = tc::kDwStages * 8 * tc::kDwSlab + 256 + 1024;
    if (!rc) launch_k(tc::k_tc_dw, ctas * n_tiles_n, tc::kDwThreads, smem, st, jobs);
  } else {
    return MARF_ERR_INVALID;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dA); cudaFree(dW); cudaFree(dAux); cudaFree(dOut); cudaFree(dBias);
  if (e != cudaSuccess) { fprintf(stderr, "marf_tc_selftest: %s\n", cudaGetErrorString(e)); return MARF_ERR_CUDA; }
  return rc;
}
