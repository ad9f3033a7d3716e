// precision=bf16: the tensor-core path.  bf16 activations/weights, fp32 accumulation in TMEM (tcgen05.mma),
// fp32 master weights and gradients.  Per layer: one k_tc_gemm launch forward, one for dX, one k_tc_dw for dW;
// encodings, the 3-/1-wide output layers, losses and reductions are SIMT kernels.  No fallback to the fp32 path.
#include <algorithm>
#include <string>

#include "engine.cuh"
#include "fp32_kernels.cuh"
#include "tc_kernels.cuh"

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
    if (e__ != cudaSuccess) return fail(h, MARF_ERR_CUDA, std::string("bf16 kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------------------------------------ SIMT helpers
__global__ void k_encode_bf16(Geo g, PxRange rg, const float* __restrict__ Hm, bf16* __restrict__ X0, int ld) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rg.padded) return;
  uint32_t* o = reinterpret_cast<uint32_t*>(X0 + (size_t)t * ld);
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
    const int L = g.L;
#pragma unroll
    for (int k = 0; k < kMaxBands; ++k) {
      if (k < L) {
        float su, cu, sv, cv;
        sincosf(u * g.band_f[k], &su, &cu);
        sincosf(v * g.band_f[k], &sv, &cv);
        float wk = g.band_w[k];
        // 2+4L <= 64 for L <= 15
        if (2 + 3 * L + k < 64) { f[2 + k] = su * wk; f[2 + L + k] = cu * wk; f[2 + 2 * L + k] = sv * wk; f[2 + 3 * L + k] = cv * wk; }
      }
    }
  }
  for (int j = 0; j < ld / 2; ++j) o[j] = j < 32 ? tc::pack_bf16(f[2 * j], f[2 * j + 1]) : 0u;
}

// mask-head features (model/planar.py:342-349) in bf16, one block per pixel row
__global__ void k_mask_features_bf16(Geo g, PxRange rg, const float* __restrict__ rgb, const float* __restrict__ embed,
                                     int embed_dim, int n_freqs, bf16* __restrict__ F, int ld) {
  int t = blockIdx.x;
  bf16* o = F + (size_t)t * ld;
  if (t >= rg.count) {
    for (int j = threadIdx.x; j < ld; j += blockDim.x) o[j] = __float2bfloat16(0.f);
    return;
  }
  int b, r, c;
  long long i = rg.first + t;
  decode_px(g, i, b, r, c);
  long long per = (long long)g.rows * g.w;
  long long rem = i - (long long)b * per;
  int k_col = 3 * embed_dim;
  for (int j = threadIdx.x; j < k_col; j += blockDim.x) {
    int ch = j / embed_dim, e = j - ch * embed_dim;
    long long idx = (long long)rgb[((long long)b * 3 + ch) * per + rem];
    o[j] = __float2bfloat16(embed[idx * embed_dim + e]);
  }
  float x, y;
  grid_xy(g, r, c, x, y);
  int k_uv = 2 + 4 * n_freqs;
  for (int j = threadIdx.x; j < k_uv; j += blockDim.x) {
    float val;
    if (j < 2) val = j == 0 ? x : y;
    else {
      int q = j - 2, fi = q / 4, w4 = q % 4;
      float a = (float)(1 << fi) * ((w4 & 1) ? y : x);
      val = (w4 < 2) ? sinf(a) : cosf(a);
    }
    o[k_col + j] = __float2bfloat16(val);
  }
  for (int j = k_col + k_uv + threadIdx.x; j < ld; j += blockDim.x) o[j] = __float2bfloat16(0.f);
}

// W fp32 [rows, cols] -> bf16 [prow, pcol] (zero padded), optionally transposed: out[c][r]
__global__ void k_pack_bf16(const float* __restrict__ W, int rows, int cols, bf16* __restrict__ out, int prow, int pcol,
                            int transpose) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= prow * pcol) return;
  int pr = i / pcol, pc = i - pr * pcol;
  int r = transpose ? pc : pr, c = transpose ? pr : pc;
  out[i] = __float2bfloat16((r < rows && c < cols) ? W[(size_t)r * cols + c] : 0.f);
}

// thin output layer (k_out <= 4): logits[row, o] = b[o] + sum_k X[row,k] W[o,k]      (fp32 out [n, 4])
template <int OUT>
__global__ void k_thin_fwd(int n, int width, const bf16* __restrict__ X, int ld, const float* __restrict__ W,
                           const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ float sWt[];   // [OUT][width]
  for (int i = threadIdx.x; i < OUT * width; i += blockDim.x) sWt[i] = W[i];
  __syncthreads();
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < n; row += nwarps) {
    float acc[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = 0.f;
    for (int k = lane * 8; k < width; k += 256) {
      uint4 raw = *reinterpret_cast<const uint4*>(X + (size_t)row * ld + k);
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float xv = __uint_as_float(((w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu) << 16);
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o] = fmaf(xv, sWt[o * width + k + e], acc[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) acc[o] = warp_sum(acc[o]);
    if (lane == 0) {
#pragma unroll
      for (int o = 0; o < 4; ++o) out[(size_t)row * 4 + o] = o < OUT ? acc[o] + bias[o] : 0.f;
    }
  }
}

// dY_prev[row,k] = (sum_o dl[row,o] W[o,k]) * (X[row,k] > 0)     (bf16 out)
template <int OUT>
__global__ void k_thin_dx(int n, int width, const float* __restrict__ dl, const float* __restrict__ W,
                          const bf16* __restrict__ X, int ld, bf16* __restrict__ dY, int ldy) {
  extern __shared__ float sWt[];
  for (int i = threadIdx.x; i < OUT * width; i += blockDim.x) sWt[i] = W[i];
  __syncthreads();
  const int per_row = width / 8;
  long long total = (long long)n * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int row = (int)(i / per_row), k = (int)(i - (long long)row * per_row) * 8;
    float d[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) d[o] = dl[(size_t)row * 4 + o];
    uint4 raw = *reinterpret_cast<const uint4*>(X + (size_t)row * ld + k);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      uint32_t hbits = (w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < OUT; ++o) s = fmaf(d[o], sWt[o * width + k + e], s);
      f[e] = ((hbits & 0x8000u) == 0 && (hbits & 0x7FFFu) != 0) ? s : 0.f;
    }
    *reinterpret_cast<uint4*>(dY + (size_t)row * ldy + k) =
        make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]), tc::pack_bf16(f[6], f[7]));
  }
}

// dW[o,k] += sum_row dl[row,o] X[row,k] ; db[o] += sum_row dl[row,o]        (fp32 atomics, W layout [OUT, ldw])
template <int OUT>
__global__ void k_thin_dw(int n, int width, const float* __restrict__ dl, const bf16* __restrict__ X, int ld,
                          float* __restrict__ dW, int ldw, float* __restrict__ db, int rows_per_block) {
  // thread t owns 8 consecutive k (width/8 threads used per row-slice); blockDim.y slices of rows
  const int kq = threadIdx.x * 8;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
  float acc[OUT][8];
  float bacc[OUT];
#pragma unroll
  for (int o = 0; o < OUT; ++o) { bacc[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[o][e] = 0.f; }
  if (kq < width) {
    for (int row = r0 + threadIdx.y; row < r1; row += blockDim.y) {
      uint4 raw = *reinterpret_cast<const uint4*>(X + (size_t)row * ld + kq);
      const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
      float d[OUT];
#pragma unroll
      for (int o = 0; o < OUT; ++o) { d[o] = dl[(size_t)row * 4 + o]; bacc[o] += d[o]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float xv = __uint_as_float(((w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu) << 16);
#pragma unroll
        for (int o = 0; o < OUT; ++o) acc[o][e] = fmaf(d[o], xv, acc[o][e]);
      }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&dW[(size_t)o * ldw + kq + e], acc[o][e]);
      if (threadIdx.x == 0) atomicAdd(&db[o], bacc[o]);
    }
  }
}

// db[j] += sum_rows dY[row, j]  (bf16 in)
__global__ void k_colsum_bf16(int n, int width, const bf16* __restrict__ dY, int ld, float* __restrict__ db, int rows_per_block) {
  // blockDim = (width/8, R): thread (tx,ty) sums 8 columns over rows ty, ty+R, ...
  __shared__ float red[8][512 + 8];
  const int kq = threadIdx.x * 8;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(n, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int row = r0 + threadIdx.y; row < r1; row += blockDim.y) {
    uint4 raw = *reinterpret_cast<const uint4*>(dY + (size_t)row * ld + kq);
    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __uint_as_float(((w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu) << 16);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.y][kq + e] = acc[e];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float s = 0.f;
      for (int y = 0; y < (int)blockDim.y; ++y) s += red[y][kq + e];
      atomicAdd(&db[kq + e], s);
    }
  }
}

__global__ void k_bf16_to_f32(long long n, const bf16* __restrict__ in, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
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
};

struct BfChain {
  int n = 0;
  BfLayer L[MARF_MAX_LAYERS];
  bf16* act[MARF_MAX_LAYERS + 1] = {};    // act[l]: input of layer l [chunk, ld[l]]
  int ld[MARF_MAX_LAYERS + 1] = {};
  CUtensorMap tmAct128[MARF_MAX_LAYERS + 1];   // box {64,128}: GEMM A loads / epilogue stores / mask loads
  CUtensorMap tmAct64[MARF_MAX_LAYERS + 1];    // box {64,64}: dW loads
  float* logits = nullptr;                // [chunk,4] fp32
  float* dlogits = nullptr;               // [chunk,4] fp32
  Chain* f32 = nullptr;                   // padded fp32 twin (gradient accumulators, bias, packing)
  bool need_dx0 = false;
};

struct Bf16State {
  EncodeTiledFn encode = nullptr;
  BfChain img, msk;
  bf16* dY[2] = {nullptr, nullptr};       // ping-pong gradient activations [chunk, max_ld]
  CUtensorMap tmDY128[2][8], tmDY64[2][8];   // per distinct ld (index = ld/64 - 1)
  int max_ld = 0;
  float* dX0 = nullptr;                   // [chunk, 64] fp32
  int num_sms = 148;
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

static int build_bf_chain(marf_handle* h, Bf16State* S, BfChain& B, Chain& F, bool need_dx0) {
  B.n = F.n;
  B.f32 = &F;
  B.need_dx0 = need_dx0;
  for (int l = 0; l < F.n; ++l) {
    BfLayer& L = B.L[l];
    L.k_in = F.k_in[l];
    L.k_out = F.k_out[l];
    L.thin = L.k_out <= 4;
    L.kp = round64(L.k_in);
    L.np = round64(L.k_out);
    B.ld[l] = L.kp;
    if (!L.thin) {
      if (l == F.n - 1) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: the last layer must be the 3-/1-wide output layer");
      if (L.k_out != 256 && L.k_out != 128)
        return fail(h, MARF_ERR_UNSUPPORTED, "bf16: hidden widths 128 and 256 are implemented (got " + std::to_string(L.k_out) + ")");
      if (L.kp > 448) return fail(h, MARF_ERR_UNSUPPORTED, "bf16: layer input wider than 448 does not fit the resident-weight tile");
      L.Wk = (bf16*)ws_alloc(h, (size_t)L.np * L.kp * 2);
      L.Wt = (bf16*)ws_alloc(h, (size_t)L.kp * L.np * 2);
      if (!L.Wk || !L.Wt) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed");
      int nt = (L.kp > 256 && L.np > 128) ? 128 : std::min(L.np, 256);   // forward N tile rows per TMA box
      int rc = make_tmap(h, S, &L.tmWk, L.Wk, L.np, L.kp, nt);
      if (rc) return rc;
      rc = make_tmap(h, S, &L.tmWt, L.Wt, L.kp, L.np, std::min(L.kp, 256));
      if (rc) return rc;
    } else if (l != F.n - 1) {
      return fail(h, MARF_ERR_UNSUPPORTED, "bf16: thin hidden layers are not supported");
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
    if (l >= 1) S->max_ld = std::max(S->max_ld, B.ld[l]);
  }
  B.logits = (float*)ws_alloc(h, (size_t)h->chunk * 4 * sizeof(float));
  B.dlogits = (float*)ws_alloc(h, (size_t)h->chunk * 4 * sizeof(float));
  if (!B.logits || !B.dlogits) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (logits)");
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
  int rc = build_bf_chain(h, S, S->img, h->img, true);
  if (rc) return rc;
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT) {
    rc = build_bf_chain(h, S, S->msk, h->msk, false);
    if (rc) return rc;
  }
  for (int i = 0; i < 2; ++i) {
    S->dY[i] = (bf16*)ws_alloc(h, (size_t)h->chunk * S->max_ld * 2);
    if (!S->dY[i]) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (dY)");
    for (int w = 1; w * 64 <= S->max_ld && w <= 8; ++w) {
      rc = make_tmap(h, S, &S->tmDY128[i][w - 1], S->dY[i], h->chunk, w * 64, 128);
      if (rc) return rc;
      rc = make_tmap(h, S, &S->tmDY64[i][w - 1], S->dY[i], h->chunk, w * 64, 64);
      if (rc) return rc;
    }
  }
  S->dX0 = (float*)ws_alloc(h, (size_t)h->chunk * 64 * sizeof(float));
  if (!S->dX0) return fail(h, MARF_ERR_CUDA, "bf16 workspace allocation failed (dX0)");
  // opt in to the large dynamic shared memory the kernels need
  const int big = 232448;
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_BIAS_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<256, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<128, tc::EPI_RELU_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_gemm<64, tc::EPI_PLAIN_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_dw<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  BF_TRY(h, cudaFuncSetAttribute(tc::k_tc_dw<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  return MARF_OK;
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
static int launch_fwd(marf_handle* h, cudaStream_t st, BfChain& B, int l, int rows) {
  Bf16State* S = h->bf16;
  BfLayer& L = B.L[l];
  tc::GemmParams p{};
  p.n_tiles = rows / 128;
  p.k_chunks = L.kp / 64;
  p.bias = B.f32->bp[l];
  int n_tile = (L.kp > 256 && L.np > 128) ? 128 : std::min(L.np, 256);
  dim3 grid(std::min(p.n_tiles, std::max(1, S->num_sms / (L.np / n_tile))), L.np / n_tile);
  int smem = tc::gemm_smem(n_tile, p.k_chunks, true).total + 1024;
  if (n_tile == 256)
    tc::k_tc_gemm<256, tc::EPI_BIAS_RELU><<<grid, tc::kThreads, smem, st>>>(B.tmAct128[l], L.tmWk, B.tmAct128[l + 1], B.tmAct128[l + 1], p);
  else
    tc::k_tc_gemm<128, tc::EPI_BIAS_RELU><<<grid, tc::kThreads, smem, st>>>(B.tmAct128[l], L.tmWk, B.tmAct128[l + 1], B.tmAct128[l + 1], p);
  BF_LAUNCH(h);
  return MARF_OK;
}

// dY_{l-1} = (dY_l W_l) * (act[l] > 0), l >= 1;  in: dY[cur] (ld = np of layer l), out: dY[cur^1] (ld = kp of layer l)
static int launch_dx(marf_handle* h, cudaStream_t st, BfChain& B, int l, int rows, int cur) {
  Bf16State* S = h->bf16;
  BfLayer& L = B.L[l];
  tc::GemmParams p{};
  p.n_tiles = rows / 128;
  p.k_chunks = L.np / 64;                    // contraction over the layer's outputs
  int n_total = L.kp;                        // produces the layer's (padded) inputs
  int n_tile = std::min(n_total, 256);
  if (n_total % n_tile) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX: input width must tile by 256/128/64");
  dim3 grid(std::min(p.n_tiles, std::max(1, S->num_sms / (n_total / n_tile))), n_total / n_tile);
  int smem = tc::gemm_smem(n_tile, p.k_chunks, true).total + 1024;
  const CUtensorMap& tmIn = S->tmDY128[cur][L.np / 64 - 1];
  const CUtensorMap& tmOut = S->tmDY128[cur ^ 1][L.kp / 64 - 1];
  if (n_tile == 256)
    tc::k_tc_gemm<256, tc::EPI_RELU_MASK><<<grid, tc::kThreads, smem, st>>>(tmIn, L.tmWt, tmOut, B.tmAct128[l], p);
  else if (n_tile == 128)
    tc::k_tc_gemm<128, tc::EPI_RELU_MASK><<<grid, tc::kThreads, smem, st>>>(tmIn, L.tmWt, tmOut, B.tmAct128[l], p);
  else
    return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX: unsupported tile");
  BF_LAUNCH(h);
  return MARF_OK;
}

static int launch_dx0(marf_handle* h, cudaStream_t st, BfChain& B, int rows, int cur) {
  Bf16State* S = h->bf16;
  BfLayer& L = B.L[0];
  if (L.kp != 64) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dX0: encoded input must fit 64 columns");
  tc::GemmParams p{};
  p.n_tiles = rows / 128;
  p.k_chunks = L.np / 64;
  p.out_f32 = S->dX0;
  p.ld_out = 64;
  p.n_store = pad4(L.k_in);
  dim3 grid(std::min(p.n_tiles, S->num_sms), 1);
  int smem = tc::gemm_smem(64, p.k_chunks, false).total + 1024;
  const CUtensorMap& tmIn = S->tmDY128[cur][L.np / 64 - 1];
  tc::k_tc_gemm<64, tc::EPI_PLAIN_F32><<<grid, tc::kThreads, smem, st>>>(tmIn, L.tmWt, tmIn, tmIn, p);
  BF_LAUNCH(h);
  return MARF_OK;
}

// dW_l += dY_l^T act[l], db_l += colsum(dY_l);  dY_l in dY[cur] with ld = np
static int launch_dw(marf_handle* h, cudaStream_t st, BfChain& B, int l, int rows, int cur) {
  Bf16State* S = h->bf16;
  BfLayer& L = B.L[l];
  Chain& F = *B.f32;
  tc::DwParams p{};
  p.rows = rows;
  p.m_halves = (L.k_out + 127) / 128;
  p.m_valid = L.k_out;
  p.n_valid = L.k_in;
  p.dW = F.gWp[l];
  p.ld_w = F.ld_in[l];
  int n_tile = L.kp >= 256 ? 256 : 64;
  if (n_tile == 64 && L.kp != 64) return fail(h, MARF_ERR_UNSUPPORTED, "bf16 dW: input width must be 64 or >= 256");
  int n_tiles_n = (L.kp + n_tile - 1) / n_tile;
  int ctas = std::max(1, S->num_sms / n_tiles_n);
  int per = (int)round_up((rows + ctas - 1) / ctas, 64);
  p.rows_per_cta = std::max(per, 64);
  ctas = (rows + p.rows_per_cta - 1) / p.rows_per_cta;
  dim3 grid(ctas, n_tiles_n);
  int stage = (p.m_halves * 2 + n_tile / 64) * tc::kDwSlab;
  int smem = tc::kDwStages * stage + 256 + 1024;
  const CUtensorMap& tmDY = S->tmDY64[cur][L.np / 64 - 1];
  if (n_tile == 256) tc::k_tc_dw<256><<<grid, tc::kThreads, smem, st>>>(tmDY, B.tmAct64[l], p);
  else tc::k_tc_dw<64><<<grid, tc::kThreads, smem, st>>>(tmDY, B.tmAct64[l], p);
  BF_LAUNCH(h);
  {
    int rpb = std::max(512, (rows + 295) / 296);
    dim3 block(L.np / 8, std::min(8, std::max(1, 256 / (L.np / 8))));
    k_colsum_bf16<<<(rows + rpb - 1) / rpb, block, 0, st>>>(rows, L.np, S->dY[cur], L.np, F.gbp[l], rpb);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

static int pack_bf_chain(marf_handle* h, cudaStream_t st, BfChain& B, const float* const* W) {
  for (int l = 0; l < B.n; ++l) {
    BfLayer& L = B.L[l];
    if (L.thin) continue;
    int tot = L.np * L.kp;
    k_pack_bf16<<<(tot + 255) / 256, 256, 0, st>>>(W[l], L.k_out, L.k_in, L.Wk, L.np, L.kp, 0);
    BF_LAUNCH(h);
    k_pack_bf16<<<(tot + 255) / 256, 256, 0, st>>>(W[l], L.k_out, L.k_in, L.Wt, L.kp, L.np, 1);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

static int thin_fwd(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* W, const float* bias) {
  int l = B.n - 1;
  int width = B.L[l].k_in;
  int out = B.L[l].k_out;
  int smem = out * width * sizeof(float);
  int blocks = std::min((rows + 7) / 8, h->bf16->num_sms * 8);
  if (out == 3) k_thin_fwd<3><<<blocks, 256, smem, st>>>(rows, width, B.act[l], B.ld[l], W, bias, B.logits);
  else if (out == 1) k_thin_fwd<1><<<blocks, 256, smem, st>>>(rows, width, B.act[l], B.ld[l], W, bias, B.logits);
  else return fail(h, MARF_ERR_UNSUPPORTED, "bf16: output layer must be 3- or 1-wide");
  BF_LAUNCH(h);
  return MARF_OK;
}

// last layer backward: dW/db (fp32 atomics into the padded fp32 twin) and dY of the previous layer -> dY[0]
static int thin_bwd(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* W) {
  Bf16State* S = h->bf16;
  Chain& F = *B.f32;
  int l = B.n - 1;
  int width = B.L[l].k_in, out = B.L[l].k_out;
  int smem = out * width * sizeof(float);
  int rpb = std::max(256, (rows + 591) / 592);
  dim3 blk(width / 8, std::max(1, 256 / (width / 8)));
  int nblk = (rows + rpb - 1) / rpb;
  int ldprev = B.ld[l];
  long long tot = (long long)rows * (width / 8);
  int dxblocks = (int)std::min<long long>((tot + 255) / 256, (long long)S->num_sms * 16);
  if (out == 3) {
    k_thin_dw<3><<<nblk, blk, 0, st>>>(rows, width, B.dlogits, B.act[l], B.ld[l], F.gWp[l], F.ld_in[l], F.gbp[l], rpb);
    BF_LAUNCH(h);
    k_thin_dx<3><<<dxblocks, 256, smem, st>>>(rows, width, B.dlogits, W, B.act[l], B.ld[l], S->dY[0], ldprev);
  } else {
    k_thin_dw<1><<<nblk, blk, 0, st>>>(rows, width, B.dlogits, B.act[l], B.ld[l], F.gWp[l], F.ld_in[l], F.gbp[l], rpb);
    BF_LAUNCH(h);
    k_thin_dx<1><<<dxblocks, 256, smem, st>>>(rows, width, B.dlogits, W, B.act[l], B.ld[l], S->dY[0], ldprev);
  }
  BF_LAUNCH(h);
  return MARF_OK;
}

static int bf_chain_forward(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* Wlast, const float* blast) {
  for (int l = 0; l < B.n - 1; ++l) {
    int rc = launch_fwd(h, st, B, l, rows);
    if (rc) return rc;
  }
  return thin_fwd(h, st, B, rows, Wlast, blast);
}

static int bf_chain_backward(marf_handle* h, cudaStream_t st, BfChain& B, int rows, const float* Wlast) {
  int rc = thin_bwd(h, st, B, rows, Wlast);      // -> dY[0] holds dY of layer n-2
  if (rc) return rc;
  int cur = 0;
  for (int l = B.n - 2; l >= 0; --l) {
    rc = launch_dw(h, st, B, l, rows, cur);
    if (rc) return rc;
    if (l > 0) {
      rc = launch_dx(h, st, B, l, rows, cur);
      if (rc) return rc;
      cur ^= 1;
    } else if (B.need_dx0) {
      rc = launch_dx0(h, st, B, rows, cur);
      if (rc) return rc;
    }
  }
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
  k_encode_bf16<<<(rg.padded + 127) / 128, 128, 0, st>>>(h->geo, rg, h->Hm, S->img.act[0], S->img.ld[0]);
  BF_LAUNCH(h);
  int rc = bf_chain_forward(h, st, S->img, rg.padded, io->mlp_w[c.n_layers - 1], io->mlp_b[c.n_layers - 1]);
  if (rc) return rc;
  if (implicit) {
    if (!(h->feats_valid && h->n_chunks == 1)) {
      k_mask_features_bf16<<<rg.padded, 128, 0, st>>>(h->geo, rg, io->rgb, io->embed, c.mask_embed_dim, c.mask_uv_freqs,
                                                      S->msk.act[0], S->msk.ld[0]);
      BF_LAUNCH(h);
      h->feats_valid = h->n_chunks == 1;
    }
    rc = bf_chain_forward(h, st, S->msk, rg.padded, io->mask_w[c.mask_n_layers - 1], io->mask_b[c.mask_n_layers - 1]);
    if (rc) return rc;
  }
  if (stats) {
    LossArgs a;
    a.mask_mode = c.mask_mode;
    a.logits = S->img.logits; a.ld = 4;
    a.mlogits = implicit ? S->msk.logits : nullptr; a.mld = 4;
    a.rgb = io->rgb; a.masks = io->masks;
    a.rgb_pred = io->rgb_pred ? io->rgb_pred : h->pred_rgb;
    a.mask_pred = io->mask_pred ? io->mask_pred : h->pred_mask;
    k_loss_stats<<<(rg.padded + 127) / 128, 128, 0, st>>>(h->geo, rg, a, io->loss_sums);
    BF_LAUNCH(h);
  }
  return MARF_OK;
}

static int bf_backward_chunk(marf_handle* h, const marf_step_io* io, cudaStream_t st, int ci) {
  Bf16State* S = h->bf16;
  const marf_config& c = h->cfg;
  PxRange rg = bf_chunk(h, ci);
  bool implicit = c.mask_mode == MARF_MASK_IMPLICIT;
  GradArgs ga;
  ga.l.mask_mode = c.mask_mode;
  ga.l.logits = S->img.logits; ga.l.ld = 4;
  ga.l.mlogits = implicit ? S->msk.logits : nullptr; ga.l.mld = 4;
  ga.l.rgb = io->rgb; ga.l.masks = io->masks; ga.l.rgb_pred = nullptr; ga.l.mask_pred = nullptr;
  ga.c_rgb = io->c_rgb; ga.c_mask = io->c_mask; ga.c_edge = io->c_edge;
  ga.edge_pred = (implicit && c.use_edges) ? (io->edge_pred ? io->edge_pred : h->edge_pred) : nullptr;
  ga.edge_label = io->edges; ga.label_channels = c.edge_label_channels > 0 ? c.edge_label_channels : 1;
  ga.dlogits = S->img.dlogits; ga.dld = 4;
  ga.dmlogits = implicit ? S->msk.dlogits : nullptr; ga.dmld = 4;
  k_loss_grad<<<(rg.padded + 127) / 128, 128, 0, st>>>(h->geo, rg, ga, h->coef);
  BF_LAUNCH(h);
  int rc = bf_chain_backward(h, st, S->img, rg.padded, io->mlp_w[c.n_layers - 1]);
  if (rc) return rc;
  k_encode_backward<<<(rg.padded + 127) / 128, 128, 0, st>>>(h->geo, rg, h->Hm, S->dX0, 64, h->G);
  BF_LAUNCH(h);
  if (implicit) {
    rc = bf_chain_backward(h, st, S->msk, rg.padded, io->mask_w[c.mask_n_layers - 1]);
    if (rc) return rc;
  }
  return MARF_OK;
}

// shared with api.cu
int engine_begin_step(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int engine_edge_pass(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int engine_begin_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st);
int engine_finish_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st);

int bf16_forward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  Bf16State* S = h->bf16;
  const marf_config& c = h->cfg;
  int rc = engine_begin_step(h, io, st);          // fp32 packing (bias), H matrices, schedule, data caches
  if (rc) return rc;
  rc = pack_bf_chain(h, st, S->img, io->mlp_w);
  if (rc) return rc;
  if (c.mask_mode == MARF_MASK_IMPLICIT) {
    rc = pack_bf_chain(h, st, S->msk, io->mask_w);
    if (rc) return rc;
  }
  BF_TRY(h, cudaMemsetAsync(io->loss_sums, 0, MARF_N_SUMS * sizeof(double), st));
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    rc = bf_forward_chunk(h, io, st, ci, true);
    if (rc) return rc;
  }
  if (c.use_edges) {
    rc = engine_edge_pass(h, io, st);
    if (rc) return rc;
  }
  h->acts_valid = h->n_chunks == 1;
  return MARF_OK;
}

int bf16_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  const marf_config& c = h->cfg;
  int rc = engine_begin_backward(h, io, st);
  if (rc) return rc;
  k_loss_coef<<<1, 1, 0, st>>>(io->loss_sums, io->norm_rgb, io->norm_edge, c.use_edges, h->coef);
  BF_LAUNCH(h);
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    if (!h->acts_valid) {
      rc = bf_forward_chunk(h, io, st, ci, false);
      if (rc) return rc;
    }
    rc = bf_backward_chunk(h, io, st, ci);
    if (rc) return rc;
  }
  h->acts_valid = false;
  return engine_finish_backward(h, io, st);
}

int bf16_step(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  int rc = bf16_forward(h, io, st);
  if (rc) return rc;
  return bf16_backward(h, io, st);
}

}  // namespace marf

// ------------------------------------------------------------------------------------------------ diagnostics
// One tensor-core kernel on caller-provided fp32 device arrays (rounded to bf16 inside), fp32 result.
//   mode 0: out[rows,N] = relu(A[rows,K] W[N,K]^T + aux[N])          (k_tc_gemm, EPI_BIAS_RELU; bf16-rounded output)
//   mode 1: out[rows,N] = (A W^T) * (aux[rows,N] > 0)                 (k_tc_gemm, EPI_RELU_MASK; bf16-rounded output)
//   mode 2: out[rows,64] = A[rows,K] W[64,K]^T                        (k_tc_gemm, EPI_PLAIN_F32)
//   mode 3: out[N(out),K(in)] = A[rows,N]^T aux[rows,K]               (k_tc_dw; A = dY, aux = X)
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
  cudaFuncSetAttribute(tc::k_tc_dw<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(tc::k_tc_dw<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  auto to_bf = [&](const float* src, int r, int c) -> bf16* {
    bf16* d = nullptr;
    if (cudaMalloc(&d, (size_t)r * c * 2) != cudaSuccess) return nullptr;
    int tot = r * c;
    k_pack_bf16<<<(tot + 255) / 256, 256, 0, st>>>(src, r, c, d, r, c, 0);
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
    tc::GemmParams p{};
    p.n_tiles = rows / 128;
    p.k_chunks = K / 64;
    dim3 grid(std::min(p.n_tiles, std::max(1, 148 / (N / n_tile))), N / n_tile);
    if (mode == 2) {
      p.out_f32 = out; p.ld_out = 64; p.n_store = 64;
      int smem = tc::gemm_smem(64, p.k_chunks, false).total + 1024;
      if (!rc) tc::k_tc_gemm<64, tc::EPI_PLAIN_F32><<<grid, tc::kThreads, smem, st>>>(tA, tW, tA, tA, p);
    } else {
      if (cudaMalloc(&dOut, (size_t)rows * N * 2) != cudaSuccess) return MARF_ERR_CUDA;
      if (!rc) rc = make_tmap(&tmp, &S, &tO, dOut, rows, N, 128);
      int smem = tc::gemm_smem(n_tile, p.k_chunks, true).total + 1024;
      if (mode == 0) {
        p.bias = aux;
        if (!rc) {
          if (n_tile == 256) tc::k_tc_gemm<256, tc::EPI_BIAS_RELU><<<grid, tc::kThreads, smem, st>>>(tA, tW, tO, tO, p);
          else if (n_tile == 128) tc::k_tc_gemm<128, tc::EPI_BIAS_RELU><<<grid, tc::kThreads, smem, st>>>(tA, tW, tO, tO, p);
          else rc = MARF_ERR_INVALID;
        }
      } else {
        dAux = to_bf(aux, rows, N);
        if (!dAux) return MARF_ERR_CUDA;
        if (!rc) rc = make_tmap(&tmp, &S, &tM, dAux, rows, N, 128);
        if (!rc) {
          if (n_tile == 256) tc::k_tc_gemm<256, tc::EPI_RELU_MASK><<<grid, tc::kThreads, smem, st>>>(tA, tW, tO, tM, p);
          else if (n_tile == 128) tc::k_tc_gemm<128, tc::EPI_RELU_MASK><<<grid, tc::kThreads, smem, st>>>(tA, tW, tO, tM, p);
          else rc = MARF_ERR_INVALID;
        }
      }
      if (!rc) {
        long long tot = (long long)rows * N;
        k_bf16_to_f32<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(tot, dOut, out);
      }
    }
  } else if (mode == 3) {
    dA = to_bf(A, rows, N);       // dY [rows, out=N]
    dAux = to_bf(aux, rows, K);   // X  [rows, in=K]
    if (!dA || !dAux) return MARF_ERR_CUDA;
    rc = make_tmap(&tmp, &S, &tA, dA, rows, N, 64);
    if (!rc) rc = make_tmap(&tmp, &S, &tM, dAux, rows, K, 64);
    tc::DwParams p{};
    p.rows = rows; p.m_halves = (N + 127) / 128; p.m_valid = N; p.n_valid = K; p.dW = out; p.ld_w = K;
    int n_tile = K >= 256 ? 256 : 64;
    if (n_tile == 64 && K != 64) return MARF_ERR_INVALID;
    if (N > 256) return MARF_ERR_INVALID;
    int n_tiles_n = (K + n_tile - 1) / n_tile;
    int ctas = std::max(1, 148 / n_tiles_n);
    p.rows_per_cta = std::max((int)round_up((rows + ctas - 1) / ctas, 64), 64);
    ctas = (rows + p.rows_per_cta - 1) / p.rows_per_cta;
    cudaMemsetAsync(out, 0, (size_t)N * K * sizeof(float), st);
    int stage = (p.m_halves * 2 + n_tile / 64) * tc::kDwSlab;
    int smem = tc::kDwStages * stage + 256 + 1024;
    dim3 grid(ctas, n_tiles_n);
    if (!rc) {
      if (n_tile == 256) tc::k_tc_dw<256><<<grid, tc::kThreads, smem, st>>>(tA, tM, p);
      else tc::k_tc_dw<64><<<grid, tc::kThreads, smem, st>>>(tA, tM, p);
    }
  } else {
    return MARF_ERR_INVALID;
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  cudaFree(dA); cudaFree(dW); cudaFree(dAux); cudaFree(dOut); cudaFree(dBias);
  if (e != cudaSuccess) { fprintf(stderr, "marf_tc_selftest: %s\n", cudaGetErrorString(e)); return MARF_ERR_CUDA; }
  return rc;
}
