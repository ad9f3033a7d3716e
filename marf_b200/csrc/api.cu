// C ABI entry points (include/marf_b200.h) and the fp32 engine that sequences the kernels.
#include <math.h>
#include <string.h>

#include <algorithm>

#include "engine.cuh"
#include "fp32_kernels.cuh"
#include "tc_tf32.cuh"

using namespace marf;

static thread_local std::string g_create_err;

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(h, MARF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

#define LAUNCH_CHECK(h)                                                                        \
  do {                                                                                         \
    (h)->launches++;                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess)                                                                    \
      return fail(h, MARF_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

namespace marf {

int fail(marf_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_create_err = msg;
  return code;
}

void* ws_alloc(marf_handle* h, size_t bytes, bool zero) {
  void* p = nullptr;
  bytes = (size_t)round_up((int64_t)std::max<size_t>(bytes, 16), 256);
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  if (zero) cudaMemset(p, 0, bytes);
  h->allocs.push_back(p);
  h->ws_bytes += (int64_t)bytes;
  return p;
}

// coarse-to-fine weights, model/planar.py:462-467, evaluated in fp32 like the reference
void set_schedule(marf_handle* h, float progress) {
  const marf_config& c = h->cfg;
  for (int k = 0; k < kMaxBands; ++k) {
    h->geo.band_f[k] = ldexpf(3.14159274101257324f, k);   // f32(2^k) * f32(pi)
    float wk = 1.0f;
    if (c.c2f_enabled) {
      float alpha = (progress - c.c2f_start) / (c.c2f_end - c.c2f_start) * (float)c.L;
      float t = fminf(fmaxf(alpha - (float)k, 0.0f), 1.0f);
      wk = (1.0f - cosf(t * 3.14159274101257324f)) / 2.0f;
    }
    h->geo.band_w[k] = wk;
  }
}

}  // namespace marf

// ------------------------------------------------------------------------------------------------
// 3xTF32 path: floats of the split form of an [N, K] weight operand (big planes + small planes, rows padded to 16, K to 32)
static size_t t32_split_floats(int N, int K) { return (size_t)2 * round_up(N, 16) * round_up(K, 32); }

static int build_chain(marf_handle* h, Chain& C, int n, const int* outs, int k_in0, uint32_t skip_mask, int d_in,
                       bool need_dx0, int act_rows) {
  C.n = n;
  C.skip_mask = skip_mask;
  C.d_in = d_in;
  C.need_dx0 = need_dx0;
  C.max_ld = 0;
  size_t total_w = 0;
  for (int l = 0; l < n; ++l) {
    C.k_out[l] = outs[l];
    C.k_in[l] = l == 0 ? k_in0 : outs[l - 1];
    if (skip_mask & (1u << l)) {
      if (l == 0 || (outs[l - 1] % 4) != 0)
        return fail(h, MARF_ERR_UNSUPPORTED, "arch.skip: layer 0 or a non-multiple-of-4 feeding width is not supported");
      C.k_in[l] += d_in;
    }
    C.ld_in[l] = pad4(C.k_in[l]);
    C.ld_out[l] = pad4(C.k_out[l]);
    C.max_ld = std::max(C.max_ld, std::max(C.ld_in[l], C.ld_out[l]));
    total_w += (size_t)C.ld_out[l] * C.ld_in[l] + C.ld_out[l];
  }
  float* wbase = (float*)ws_alloc(h, total_w * sizeof(float));
  float* gbase = (float*)ws_alloc(h, total_w * sizeof(float));
  if (!wbase || !gbase) return fail(h, MARF_ERR_CUDA, "workspace allocation failed (weights)");
  size_t off = 0;
  for (int l = 0; l < n; ++l) {
    C.Wp[l] = wbase + off; C.gWp[l] = gbase + off; off += (size_t)C.ld_out[l] * C.ld_in[l];
    C.bp[l] = wbase + off; C.gbp[l] = gbase + off; off += C.ld_out[l];
  }
  for (int l = 0; l < n && h->fp32_tc && act_rows > 0; ++l) {
    if (C.ld_out[l] < 32 || C.ld_in[l] < 32) continue;
    C.Wsp_f[l] = (float*)ws_alloc(h, t32_split_floats(C.ld_out[l], C.ld_in[l]) * sizeof(float));
    C.Wt[l] = (float*)ws_alloc(h, t32_split_floats(C.ld_in[l], C.ld_out[l]) * sizeof(float));
    if (!C.Wsp_f[l] || !C.Wt[l]) return fail(h, MARF_ERR_CUDA, "workspace allocation failed (split weights)");
    if (l + 1 < n && !(skip_mask & (1u << (l + 1)))) {      // sign bits of this layer's output = the input of layer l + 1
      C.bits_ld[l + 1] = (C.ld_out[l] + 31) / 32;
      C.bits[l + 1] = (uint32_t*)ws_alloc(h, (size_t)act_rows * C.bits_ld[l + 1] * sizeof(uint32_t));
      if (!C.bits[l + 1]) return fail(h, MARF_ERR_CUDA, "workspace allocation failed (sign bits)");
    }
  }
  for (int l = 0; l <= n; ++l) {
    int ld = l < n ? C.ld_in[l] : C.ld_out[n - 1];
    C.act[l] = act_rows > 0 ? (float*)ws_alloc(h, (size_t)act_rows * ld * sizeof(float)) : (float*)ws_alloc(h, 16);
    if (!C.act[l]) return fail(h, MARF_ERR_CUDA, "workspace allocation failed (activations)");
  }
  return MARF_OK;
}

static size_t chain_param_floats(const Chain& C) {
  size_t t = 0;
  for (int l = 0; l < C.n; ++l) t += (size_t)C.ld_out[l] * C.ld_in[l] + C.ld_out[l];
  return t;
}

extern "C" int marf_abi_version(void) { return MARF_ABI_VERSION; }

extern "C" const char* marf_last_error(const marf_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int64_t marf_launch_count(const marf_handle* h) { return h ? h->launches : 0; }

extern "C" int marf_profile(marf_handle* h, int enable) {
  if (!h) return MARF_ERR_INVALID;
  h->profiling = enable != 0;
  return MARF_OK;
}

extern "C" int marf_profile_read(marf_handle* h, double* ms, int64_t* launches, int n_classes) {
  if (!h || !ms || !launches || n_classes > MARF_PROF_CLASSES) return MARF_ERR_INVALID;
  for (int c = 0; c < n_classes; ++c) {
    ms[c] = 0.0;
    launches[c] = (int64_t)h->prof_ev[c].size();
    for (auto& pr : h->prof_ev[c]) {
      if (cudaEventSynchronize(pr.second) != cudaSuccess) return marf::fail(h, MARF_ERR_CUDA, "marf_profile_read: event sync failed");
      float t = 0.f;
      if (cudaEventElapsedTime(&t, pr.first, pr.second) != cudaSuccess) return marf::fail(h, MARF_ERR_CUDA, "marf_profile_read: elapsed time failed");
      ms[c] += t;
      h->prof_pool.push_back(pr.first);
      h->prof_pool.push_back(pr.second);
    }
    h->prof_ev[c].clear();
  }
  return MARF_OK;
}
extern "C" int64_t marf_workspace_bytes(const marf_handle* h) { return h ? h->ws_bytes : 0; }

extern "C" int marf_destroy(marf_handle* h) {
  if (!h) return MARF_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  bf16_destroy(h);
  for (void* p : h->allocs) cudaFree(p);
  for (auto& v : h->prof_ev)
    for (auto& pr : v) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  delete h;
  return MARF_OK;
}

static int t32_init(marf_handle* h);      // 3xTF32 kernels: per-handle set-up (defined with the fp32 engine below)

extern "C" int marf_create(const marf_config* cfg, marf_handle** out) {
  if (!cfg || !out) return fail(nullptr, MARF_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MARF_ABI_VERSION) return fail(nullptr, MARF_ERR_INVALID, "abi_version mismatch");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(nullptr, MARF_ERR_NO_DEVICE, "no CUDA device: marf_b200 has no CPU path");
  }
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MARF_ERR_INVALID, "bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10)
    return fail(nullptr, MARF_ERR_NO_DEVICE, "device is not sm_100 (B200); this library ships sm_100a code only");
  const marf_config& c = *cfg;
  if (c.H <= 0 || c.W <= 0 || c.patch_H <= 0 || c.patch_W <= 0 || c.batch_global <= 0)
    return fail(nullptr, MARF_ERR_INVALID, "non-positive geometry");
  if (c.use_cropped && ((c.patch_H & 1) || (c.patch_W & 1) || c.patch_H > c.H || c.patch_W > c.W))
    return fail(nullptr, MARF_ERR_INVALID, "cropped patches must have even sizes <= the canvas (warp.py:14-19)");
  if (c.L < 0 || c.L > kMaxBands) return fail(nullptr, MARF_ERR_INVALID, "arch.posenc.L_2D out of range [0,16]");
  if (c.n_layers < 1 || c.n_layers > MARF_MAX_LAYERS) return fail(nullptr, MARF_ERR_INVALID, "bad n_layers");
  if (c.layer_out[c.n_layers - 1] != 3) return fail(nullptr, MARF_ERR_INVALID, "last layer must have 3 outputs (rgb)");
  if (c.mask_mode < 0 || c.mask_mode > 2) return fail(nullptr, MARF_ERR_INVALID, "bad mask_mode");
  if (c.precision != MARF_FP32 && c.precision != MARF_BF16) return fail(nullptr, MARF_ERR_INVALID, "bad precision");
  const int hh = c.use_cropped ? c.patch_H : c.H, ww = c.use_cropped ? c.patch_W : c.W;
  if (c.batch <= 0 || c.patch_offset < 0 || c.patch_offset + c.batch > c.batch_global)
    return fail(nullptr, MARF_ERR_INVALID, "bad patch shard");
  if (c.rows <= 0 || c.row_offset < 0 || c.row_offset + c.rows > hh) return fail(nullptr, MARF_ERR_INVALID, "bad row shard");
  if (c.use_edges && c.rows != hh)
    return fail(nullptr, MARF_ERR_UNSUPPORTED, "use_edges needs whole patches per shard (Sobel/Gauss halo)");
  if (c.mask_mode == MARF_MASK_IMPLICIT) {
    if (c.mask_n_layers < 1 || c.mask_n_layers > MARF_MAX_LAYERS || c.mask_layer_out[c.mask_n_layers - 1] != 1)
      return fail(nullptr, MARF_ERR_INVALID, "mask head must end in 1 output");
    if (c.mask_uv_freqs < 0 || c.mask_uv_freqs > 16 || c.mask_embed_dim <= 0)
      return fail(nullptr, MARF_ERR_INVALID, "bad mask embedding sizes");
  }
  if (cudaSetDevice(c.device) != cudaSuccess) return fail(nullptr, MARF_ERR_CUDA, "cudaSetDevice failed");

  marf_handle* h = new marf_handle();
  h->cfg = c;
  h->n_sms = prop.multiProcessorCount;
  h->fp32_tc = getenv("MARF_FP32_TC") ? atoi(getenv("MARF_FP32_TC")) != 0 : true;
  h->h = hh;
  h->w = ww;
  Geo& g = h->geo;
  g.H = c.H; g.W = c.W; g.h = hh; g.w = ww;
  g.y0 = c.use_cropped ? c.H / 2 - c.patch_H / 2 : 0;
  g.x0 = c.use_cropped ? c.W / 2 - c.patch_W / 2 : 0;
  g.rows = c.rows; g.row_offset = c.row_offset; g.patch_offset = c.patch_offset;
  g.norm_h = (float)((double)c.H / (double)std::max(c.H, c.W));
  g.norm_w = (float)((double)c.W / (double)std::max(c.H, c.W));
  g.L = c.L;
  g.d_in = c.L > 0 ? 2 + 4 * c.L : 2;
  set_schedule(h, 0.f);
  h->n_local = (long long)c.batch * c.rows * ww;
  long long max_chunk = c.max_chunk_pixels > 0 ? c.max_chunk_pixels : (c.precision == MARF_BF16 ? (1ll << 20) : (1ll << 19));
  // the render path traverses up to H*W pixels per patch with the same buffers
  h->chunk = (int)round_up(std::min<long long>(std::max<long long>(h->n_local, 128), max_chunk), 128);
  h->n_chunks = (int)((h->n_local + h->chunk - 1) / h->chunk);
  // fp32 activation buffers: the whole chunk in fp32 mode; in bf16 mode only the forward-only render uses them
  const bool f32 = c.precision == MARF_FP32;
  h->render_rows = f32 ? h->chunk : std::min(h->chunk, 1 << 16);

  int rc = build_chain(h, h->img, c.n_layers, c.layer_out, g.d_in, c.skip_mask, g.d_in, true, h->render_rows);
  if (rc == MARF_OK && c.mask_mode == MARF_MASK_IMPLICIT)
    rc = build_chain(h, h->msk, c.mask_n_layers, c.mask_layer_out, 3 * c.mask_embed_dim + 2 + 4 * c.mask_uv_freqs, 0, 0, false,
                     f32 ? h->chunk : 0);
  if (rc != MARF_OK) { std::string e = h->err; marf_destroy(h); g_create_err = e; return rc; }
  int max_ld = std::max(h->img.max_ld, c.mask_mode == MARF_MASK_IMPLICIT ? h->msk.max_ld : 0);
  h->Hm = (float*)ws_alloc(h, (size_t)c.batch_global * 9 * sizeof(float));
  h->G = (double*)ws_alloc(h, (size_t)c.batch * 9 * sizeof(double));
  h->dYa = (float*)ws_alloc(h, f32 ? (size_t)h->chunk * max_ld * sizeof(float) : 16);
  h->dYb = (float*)ws_alloc(h, f32 ? (size_t)h->chunk * max_ld * sizeof(float) : 16);
  h->coef = (LossCoef*)ws_alloc(h, sizeof(LossCoef));
  h->sums_static = (double*)ws_alloc(h, 4 * sizeof(double));
  h->bad_index = (double*)ws_alloc(h, sizeof(double));
  if (h->cfg.mask_n_vocab <= 0) h->cfg.mask_n_vocab = 1500;        // options/planar.yaml N_vocab
  bool ok = h->Hm && h->G && h->dYa && h->dYb && h->coef && h->sums_static && h->bad_index;
  if (c.skip_mask) {
    h->dX0acc = (float*)ws_alloc(h, (size_t)h->chunk * h->img.ld_in[0] * sizeof(float));
    h->dXscratch = (float*)ws_alloc(h, (size_t)h->chunk * max_ld * sizeof(float));
    ok = ok && h->dX0acc && h->dXscratch;
  }
  if (c.use_edges) {
    h->pred_rgb = (float*)ws_alloc(h, (size_t)h->n_local * 3 * sizeof(float));
    h->pred_mask = (float*)ws_alloc(h, (size_t)h->n_local * sizeof(float));
    h->edge_mag = (double*)ws_alloc(h, (size_t)h->n_local * 3 * sizeof(double));
    h->edge_pred = (double*)ws_alloc(h, (size_t)h->n_local * 3 * sizeof(double));
    ok = ok && h->pred_rgb && h->pred_mask && h->edge_mag && h->edge_pred;
  }
  if (!ok) { marf_destroy(h); return fail(nullptr, MARF_ERR_CUDA, "workspace allocation failed"); }
  if (h->fp32_tc) {
    rc = t32_init(h);
    if (rc != MARF_OK) { std::string e = h->err; marf_destroy(h); g_create_err = e; return rc; }
  }
  if (c.precision == MARF_BF16) {
    rc = bf16_create(h);
    if (rc != MARF_OK) { std::string e = h->err; marf_destroy(h); g_create_err = e; return rc; }
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { marf_destroy(h); return fail(nullptr, MARF_ERR_CUDA, "create: device sync failed"); }
  *out = h;
  return MARF_OK;
}

// ------------------------------------------------------------------------------------------------
// fp32 engine
// ------------------------------------------------------------------------------------------------
template <bool A_KC, bool B_KC, int EPI>
static int sgemm(marf_handle* h, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 float* C, int ldc, const float* aux, int ldaux, int k_split) {
  if (M <= 0 || N <= 0 || K <= 0) return MARF_OK;
  if (N <= 16) {
    dim3 grid((M + GBM - 1) / GBM, (N + 15) / 16, k_split);
    launch_k(k_sgemm<A_KC, B_KC, 1, EPI>, grid, 256, 0, st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, k_split);
  } else {
    dim3 grid((M + GBM - 1) / GBM, (N + 127) / 128, k_split);
    launch_k(k_sgemm<A_KC, B_KC, 8, EPI>, grid, 256, 0, st, M, N, K, A, lda, B, ldb, C, ldc, aux, ldaux, k_split);
  }
  LAUNCH_CHECK(h);
  return MARF_OK;
}

// ---- 3xTF32 tensor-core GEMMs (tc_tf32.cuh)
// C[M, N] = epi(A[M, K] * B[N, K]^T), B row-major [N, ldb]
// fp32 [rows, cols] tensor map, box [32 rows x 32 columns], SWIZZLE_128B, out-of-range columns read as zero
static int t32_tmap(marf_handle* h, CUtensorMap* m, const float* base, int rows, int cols, int ld, int box_rows = 32) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
      return fail(h, MARF_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    encode = (EncodeFn)fn;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, MARF_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed (code " + std::to_string((int)r) + ")");
  return MARF_OK;
}

// once per handle (marf_create): dynamic-SMEM opt-in of every 3xTF32 kernel on this device, and the tensor-map encoder — so that
// the first step does nothing but launches (it may be captured into a CUDA graph)
template <int MODE, int EPI>
static cudaError_t t32_optin() {
  return cudaFuncSetAttribute(t32::k_tf32x3<MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, t32::kSmemBytes);
}
static int t32_init(marf_handle* h) {
  cudaError_t e = t32_optin<t32::MODE_NT, t32::T_BIAS>();
  if (e == cudaSuccess) e = t32_optin<t32::MODE_NT, t32::T_BIAS_RELU>();
  if (e == cudaSuccess) e = t32_optin<t32::MODE_NT, t32::T_PLAIN>();
  if (e == cudaSuccess) e = t32_optin<t32::MODE_NT, t32::T_RELU_MASK>();
  if (e == cudaSuccess) e = t32_optin<t32::MODE_NT, t32::T_RELU_BITS>();
  if (e == cudaSuccess) e = t32_optin<t32::MODE_TN, t32::T_PLAIN>();
  if (e != cudaSuccess) return fail(h, MARF_ERR_CUDA, std::string("3xTF32 kernels: ") + cudaGetErrorString(e));
  CUtensorMap probe;
  return t32_tmap(h, &probe, reinterpret_cast<const float*>(h->Hm), 32, 32, 32);   // resolves the driver entry point
}

static long long* g_t32_trace = nullptr;      // diagnostics only (marf_tf32_gemm with MARF_T32_TRACE=1)

// weights -> split planes (k_tf32_split_w)
template <int TRANS>
static int tf32_split(marf_handle* h, cudaStream_t st, const float* W, int ldw, int N, int K, float* out) {
  const int nr = (int)round_up(N, 16), kpad = (int)round_up(K, 32);
  launch_k(t32::k_tf32_split_w<TRANS>, (nr * kpad + 255) / 256, 256, 0, st, W, ldw, N, K, out, nr, kpad);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

// C[M, N] = epi(A[M, K] * B[N, K]^T), B given in split form (tf32_split)
template <int EPI>
static int tgemm_nt(marf_handle* h, cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* Bsp, float* C,
                    int ldc, const float* aux, int ldaux, uint32_t* bits = nullptr, int bits_ld = 0) {
  const int nr = (int)round_up(N, 16), kpad = (int)round_up(K, 32);
  for (int n0 = 0; n0 < N; n0 += 256) {
    t32::Params p{};
    p.A = A; p.lda = lda;
    p.C = C + n0; p.ldc = ldc;
    p.aux = aux ? aux + n0 : nullptr; p.ldaux = ldaux;
    p.M = M; p.K = K;
    p.n_valid = std::min(256, N - n0);
    if (bits) {
      if (EPI == t32::T_RELU_BITS) p.bits_in = bits + n0 / 32; else p.bits_out = bits + n0 / 32;
      p.bits_ld = bits_ld;
    }
    p.trace = g_t32_trace;
    p.b_row0 = n0;
    p.b_small = nr;
    int rc = t32_tmap(h, &p.tmB, Bsp, 2 * nr, kpad, kpad, (int)round_up(p.n_valid, 16) / 2);
    if (rc == MARF_OK && EPI != t32::T_RELU_MASK) rc = t32_tmap(h, &p.tmC, C + n0, M, p.n_valid, ldc);
    if (rc) return rc;
    const int pairs = std::min((M / t32::kTileM + 1) / 2, h->n_sms / 2);
    launch_k_cluster(t32::k_tf32x3<t32::MODE_NT, EPI>, 2 * pairs, t32::kThreads, t32::kSmemBytes, st, 2, p);
    LAUNCH_CHECK(h);
  }
  return MARF_OK;
}

// C[Np, Nq] += P[M, Np]^T * Q[M, Nq];  db[Np] += column sums of P (optional)
static int tgemm_tn(marf_handle* h, cudaStream_t st, int M, int Np, int Nq, const float* P, int ldp, const float* Q, int ldq, float* C,
                    int ldc, float* db) {
  t32::Params p{};
  p.A = P; p.lda = ldp; p.B = Q; p.ldb = ldq;
  p.C = C; p.ldc = ldc;
  p.db = db;
  p.trace = g_t32_trace;
  int rc = t32_tmap(h, &p.tmP, P, M, Np, ldp);
  if (rc == MARF_OK) rc = t32_tmap(h, &p.tmQ, Q, M, Nq, ldq);
  if (rc) return rc;
  p.M = M;
  p.p_valid = Np; p.n_valid = Nq;
  dim3 grid(1, (Np + 255) / 256, (Nq + 255) / 256);
  p.splits = std::max(1, std::min(M / t32::kStageK, h->n_sms / 2 / (int)(grid.y * grid.z)));
  grid.x = 2 * p.splits;
  launch_k_cluster(t32::k_tf32x3<t32::MODE_TN, t32::T_PLAIN>, grid, t32::kThreadsTN, t32::kSmemBytes, st, 2, p);
  LAUNCH_CHECK(h);
  return MARF_OK;
}
// the narrow output layer as bandwidth-bound row kernels (fp32_kernels.cuh: k_out_forward / k_out_backward)
static bool out_rows_ok(const marf_handle* h, const Chain& C, int l) {
  return h->fp32_tc && C.ld_out[l] == 4 && C.ld_in[l] <= 512 && C.ld_in[l] % 4 == 0;
}
static int out_forward(marf_handle* h, cudaStream_t st, Chain& C, int l, int M) {
  const int K = C.ld_in[l], kch = (K + 127) / 128, grid = std::min((M + 7) / 8, 8 * h->n_sms);
#define MARF_OUT_FWD(KCH)                                                                                                       \
  launch_k(k_out_forward<KCH>, grid, 256, 0, st, M, K, C.k_out[l], C.act[l], C.ld_in[l], C.Wp[l], C.ld_in[l], C.bp[l], C.act[l + 1], \
           C.ld_out[l])
  if (kch == 1) MARF_OUT_FWD(1); else if (kch == 2) MARF_OUT_FWD(2); else if (kch == 3) MARF_OUT_FWD(3); else MARF_OUT_FWD(4);
#undef MARF_OUT_FWD
  LAUNCH_CHECK(h);
  return MARF_OK;
}
static int out_backward(marf_handle* h, cudaStream_t st, Chain& C, int l, int M, const float* dY, float* dX) {
  const int K = C.ld_in[l], kch = (K + 127) / 128, grid = std::min((M + 7) / 8, 4 * h->n_sms);
#define MARF_OUT_BWD(KCH)                                                                                                        \
  launch_k(k_out_backward<KCH>, grid, 256, 0, st, M, K, C.k_out[l], C.act[l], C.ld_in[l], dY, C.ld_out[l], C.Wp[l], C.ld_in[l], dX, \
           C.ld_in[l], C.gWp[l], C.gbp[l])
  if (kch == 1) MARF_OUT_BWD(1); else if (kch == 2) MARF_OUT_BWD(2); else if (kch == 3) MARF_OUT_BWD(3); else MARF_OUT_BWD(4);
#undef MARF_OUT_BWD
  LAUNCH_CHECK(h);
  return MARF_OK;
}
static bool tc_rows_ok(const marf_handle* h, int M) { return h->fp32_tc && M >= 128 && M % 128 == 0; }

extern "C" int marf_tf32_gemm(marf_handle* h, int mode, int epi, int M, int N, int K, const float* A, int lda, const float* W, int ldw,
                              float* C, int ldc, const float* aux, int ldaux, void* stream) {
  if (!h || !A || !W || !C) return MARF_ERR_INVALID;
  if (M < 128 || M % 128 || (lda | ldw | ldc | ldaux) % 4 || mode < 0 || mode > 2 || epi < 0 || epi > 3)
    return fail(h, MARF_ERR_INVALID, "marf_tf32_gemm: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = MARF_OK;
  float* wt = nullptr;
  const bool trace = getenv("MARF_T32_TRACE") != nullptr;
  if (trace) {
    CUDA_TRY(h, cudaMalloc(&g_t32_trace, 512 * 8 * sizeof(long long)));
    CUDA_TRY(h, cudaMemset(g_t32_trace, 0, 512 * 8 * sizeof(long long)));
  }
  if (mode == 2) {
    rc = tgemm_tn(h, st, M, N, K, A, lda, W, ldw, C, ldc, const_cast<float*>(aux));
  } else {
    CUDA_TRY(h, cudaMalloc(&wt, t32_split_floats(N, K) * sizeof(float)));
    rc = mode == 0 ? tf32_split<0>(h, st, W, ldw, N, K, wt) : tf32_split<1>(h, st, W, ldw, N, K, wt);
    if (rc == MARF_OK)
      rc = epi == 0 ? tgemm_nt<t32::T_BIAS>(h, st, M, N, K, A, lda, wt, C, ldc, aux, ldaux)
         : epi == 1 ? tgemm_nt<t32::T_BIAS_RELU>(h, st, M, N, K, A, lda, wt, C, ldc, aux, ldaux)
         : epi == 2 ? tgemm_nt<t32::T_PLAIN>(h, st, M, N, K, A, lda, wt, C, ldc, aux, ldaux)
                    : tgemm_nt<t32::T_RELU_MASK>(h, st, M, N, K, A, lda, wt, C, ldc, aux, ldaux);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (trace && e == cudaSuccess) {
    std::vector<long long> t(512 * 8);
    cudaMemcpy(t.data(), g_t32_trace, t.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    const long long t0 = t[0];
    fprintf(stderr, "# stage: loader[empty seen, stored, arrived] mma[full seen, committed]   (cycles since the first stamp)\n");
    for (int i = 0; i < 40; ++i)
      fprintf(stderr, "%3d: L %7lld %7lld %7lld   M %7lld %7lld   (dW form: last loader warp %7lld, peer CTA %7lld)\n", i, t[i * 8] - t0,
              t[i * 8 + 1] - t0, t[i * 8 + 2] - t0, t[i * 8 + 4] - t0, t[i * 8 + 5] - t0, t[i * 8 + 3] - t0, t[i * 8 + 6] - t0);
    fprintf(stderr, "# tile: epilogue[acc_full seen, released]\n");
    for (int i = 0; i < 6; ++i) fprintf(stderr, "%3d: E %7lld %7lld\n", i, t[i * 8 + 6] - t0, t[i * 8 + 7] - t0);
    cudaFree(g_t32_trace);
    g_t32_trace = nullptr;
  }
  if (wt) cudaFree(wt);
  if (e != cudaSuccess) return fail(h, MARF_ERR_CUDA, std::string("marf_tf32_gemm: ") + cudaGetErrorString(e));
  return rc;
}

static int pick_split(int M, int N, int K) {
  long long tiles = (long long)((M + GBM - 1) / GBM) * ((N + (N <= 16 ? 15 : 127)) / (N <= 16 ? 16 : 128));
  int kt = (K + GBK - 1) / GBK;
  long long want = (592 + tiles - 1) / tiles;       // ~4 CTAs per SM
  return (int)std::max<long long>(1, std::min<long long>(want, std::max(1, kt / 8)));
}

static int pack_chain(marf_handle* h, cudaStream_t st, Chain& C, const float* const* W, const float* const* b) {
  if (!W || !b) return fail(h, MARF_ERR_INVALID, "missing parameter pointers");
  ChainPackJobs J{};
  int biggest = 0;
  for (int l = 0; l < C.n; ++l) {
    if (!W[l] || !b[l]) return fail(h, MARF_ERR_INVALID, "null layer parameter");
    J.W[l] = W[l]; J.b[l] = b[l]; J.Wp[l] = C.Wp[l]; J.bp[l] = C.bp[l];
    J.sp_f[l] = C.Wt[l] ? C.Wsp_f[l] : nullptr;
    J.sp_t[l] = C.Wt[l];
    J.k_out[l] = C.k_out[l]; J.k_in[l] = C.k_in[l]; J.ld_out[l] = C.ld_out[l]; J.ld_in[l] = C.ld_in[l];
    biggest = std::max(biggest, (int)round_up(C.ld_out[l], 16) * (int)round_up(C.ld_in[l], 32));
  }
  launch_k(k_pack_chain, dim3(std::max(1, std::min(64, (biggest + 1023) / 1024)), C.n), 256, 0, st, J);
  LAUNCH_CHECK(h);
  return MARF_OK;
}
static int unpack_chain(marf_handle* h, cudaStream_t st, Chain& C, float* const* gW, float* const* gb) {
  if (!gW || !gb) return fail(h, MARF_ERR_INVALID, "missing gradient pointers");
  ChainUnpackJobs J{};
  int biggest = 0;
  for (int l = 0; l < C.n; ++l) {
    J.gWp[l] = C.gWp[l]; J.gbp[l] = C.gbp[l]; J.gW[l] = gW[l]; J.gb[l] = gb[l];
    J.k_out[l] = C.k_out[l]; J.k_in[l] = C.k_in[l]; J.ld_in[l] = C.ld_in[l];
    biggest = std::max(biggest, C.k_out[l] * C.k_in[l]);
  }
  launch_k(k_unpack_chain, dim3(std::max(1, std::min(64, (biggest + 1023) / 1024)), C.n), 256, 0, st, J);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

static int chain_forward(marf_handle* h, cudaStream_t st, Chain& C, int M) {
  for (int l = 0; l < C.n; ++l) {
    bool last = l == C.n - 1;
    int ldc = last ? C.ld_out[l] : C.ld_in[l + 1];
    int rc;
    if (!last) C.bits_ok[l + 1] = false;
    if (last && out_rows_ok(h, C, l)) {
      rc = out_forward(h, st, C, l, M);
    } else if (C.Wt[l] && tc_rows_ok(h, M)) {
      if (!last) C.bits_ok[l + 1] = C.bits[l + 1] != nullptr;
      rc = last ? tgemm_nt<t32::T_BIAS>(h, st, M, C.ld_out[l], C.ld_in[l], C.act[l], C.ld_in[l], C.Wsp_f[l], C.act[l + 1], ldc, C.bp[l], 0)
                : tgemm_nt<t32::T_BIAS_RELU>(h, st, M, C.ld_out[l], C.ld_in[l], C.act[l], C.ld_in[l], C.Wsp_f[l], C.act[l + 1],
                                             ldc, C.bp[l], 0, C.bits[l + 1], C.bits_ld[l + 1]);
    } else
      rc = last ? sgemm<true, true, EPI_BIAS>(h, st, M, C.ld_out[l], C.ld_in[l], C.act[l], C.ld_in[l], C.Wp[l], C.ld_in[l],
                                              C.act[l + 1], ldc, C.bp[l], 0, 1)
                : sgemm<true, true, EPI_BIAS_RELU>(h, st, M, C.ld_out[l], C.ld_in[l], C.act[l], C.ld_in[l], C.Wp[l],
                                                   C.ld_in[l], C.act[l + 1], ldc, C.bp[l], 0, 1);
    if (rc) return rc;
    if (!last && (C.skip_mask & (1u << (l + 1)))) {
      long long tot = (long long)M * C.d_in;
      launch_k(k_copy_cols, (unsigned)((tot + 255) / 256), 256, 0, st, M, C.d_in, C.act[0], C.ld_in[0], 0, C.act[l + 1],
                                                                 C.ld_in[l + 1], C.k_out[l], 0);
      LAUNCH_CHECK(h);
    }
  }
  return MARF_OK;
}

// dY of the last layer must be in h->dYa [M, ld_out[n-1]].  Returns the buffer holding dX0 in *dx0 (if need_dx0).
static int chain_backward(marf_handle* h, cudaStream_t st, Chain& C, int M, float** dx0) {
  float* cur = h->dYa;
  float* nxt = h->dYb;
  if (C.skip_mask) CUDA_TRY(h, cudaMemsetAsync(h->dX0acc, 0, (size_t)M * C.ld_in[0] * sizeof(float), st));
  for (int l = C.n - 1; l >= 0; --l) {
    int ldy = C.ld_out[l];
    int rc;
    const bool tcl = C.Wt[l] && tc_rows_ok(h, M);
    if (l == C.n - 1 && l > 0 && out_rows_ok(h, C, l) && !(C.skip_mask & (1u << l))) {
      // output layer: dW, db and the masked dX in one pass over the layer input
      rc = out_backward(h, st, C, l, M, cur, nxt);
      if (rc) return rc;
      std::swap(cur, nxt);
      continue;
    }
    if (tcl) {
      rc = tgemm_tn(h, st, M, C.ld_out[l], C.ld_in[l], cur, ldy, C.act[l], C.ld_in[l], C.gWp[l], C.ld_in[l], C.gbp[l]);
    } else if (C.ld_out[l] <= 16) {
      rc = sgemm<false, false, EPI_ATOMIC_T>(h, st, C.ld_in[l], C.ld_out[l], M, C.act[l], C.ld_in[l], cur, ldy, C.gWp[l],
                                             C.ld_in[l], nullptr, 0, pick_split(C.ld_in[l], C.ld_out[l], M));
    } else {
      rc = sgemm<false, false, EPI_ATOMIC>(h, st, C.ld_out[l], C.ld_in[l], M, cur, ldy, C.act[l], C.ld_in[l], C.gWp[l],
                                           C.ld_in[l], nullptr, 0, pick_split(C.ld_out[l], C.ld_in[l], M));
    }
    if (rc) return rc;
    if (!tcl) {                 // (the tensor-core dW kernel sums the columns of dY while it splits them)
      int rpb = std::max(256, (M + 63) / 64);
      dim3 grid((C.ld_out[l] + 31) / 32, (M + rpb - 1) / rpb);
      launch_k(k_colsum, grid, 256, 0, st, M, C.ld_out[l], cur, ldy, C.gbp[l], rpb);
      LAUNCH_CHECK(h);
    }
    if (l == 0 && !C.need_dx0) break;
    if (l == 0) {
      rc = tcl ? tgemm_nt<t32::T_PLAIN>(h, st, M, C.ld_in[0], C.ld_out[0], cur, ldy, C.Wt[0], nxt, C.ld_in[0], nullptr, 0)
               : sgemm<true, false, EPI_PLAIN>(h, st, M, C.ld_in[0], C.ld_out[0], cur, ldy, C.Wp[0], C.ld_in[0], nxt, C.ld_in[0],
                                               nullptr, 0, 1);
      if (rc) return rc;
      if (C.skip_mask) {
        long long tot = (long long)M * C.d_in;
        launch_k(k_copy_cols, (unsigned)((tot + 255) / 256), 256, 0, st, M, C.d_in, h->dX0acc, C.ld_in[0], 0, nxt, C.ld_in[0], 0, 1);
        LAUNCH_CHECK(h);
      }
    } else if (C.skip_mask & (1u << l)) {
      rc = tcl ? tgemm_nt<t32::T_PLAIN>(h, st, M, C.ld_in[l], C.ld_out[l], cur, ldy, C.Wt[l], h->dXscratch, C.ld_in[l], nullptr, 0)
               : sgemm<true, false, EPI_PLAIN>(h, st, M, C.ld_in[l], C.ld_out[l], cur, ldy, C.Wp[l], C.ld_in[l], h->dXscratch,
                                               C.ld_in[l], nullptr, 0, 1);
      if (rc) return rc;
      long long tot = (long long)M * C.k_out[l - 1];
      launch_k(k_relu_mask, (unsigned)((tot + 255) / 256), 256, 0, st, M, C.k_out[l - 1], h->dXscratch, C.ld_in[l], C.act[l],
                                                                 C.ld_in[l], nxt, C.ld_out[l - 1]);
      LAUNCH_CHECK(h);
      tot = (long long)M * C.d_in;
      launch_k(k_copy_cols, (unsigned)((tot + 255) / 256), 256, 0, st, M, C.d_in, h->dXscratch, C.ld_in[l], C.k_out[l - 1],
                                                                 h->dX0acc, C.ld_in[0], 0, 1);
      LAUNCH_CHECK(h);
    } else {
      rc = tcl && C.bits_ok[l]
               ? tgemm_nt<t32::T_RELU_BITS>(h, st, M, C.ld_in[l], C.ld_out[l], cur, ldy, C.Wt[l], nxt, C.ld_in[l], nullptr, 0,
                                            C.bits[l], C.bits_ld[l])
           : tcl ? tgemm_nt<t32::T_RELU_MASK>(h, st, M, C.ld_in[l], C.ld_out[l], cur, ldy, C.Wt[l], nxt, C.ld_in[l], C.act[l],
                                            C.ld_in[l])
               : sgemm<true, false, EPI_RELU_MASK>(h, st, M, C.ld_in[l], C.ld_out[l], cur, ldy, C.Wp[l], C.ld_in[l], nxt,
                                                   C.ld_in[l], C.act[l], C.ld_in[l], 1);
      if (rc) return rc;
    }
    std::swap(cur, nxt);
  }
  if (dx0) *dx0 = cur;
  // keep the invariant "dY of the last layer is written to dYa" for the next chain: callers always refill dYa
  return MARF_OK;
}

static PxRange chunk_range(const marf_handle* h, int ci) {
  PxRange rg;
  rg.first = (long long)ci * h->chunk;
  rg.count = (int)std::min<long long>(h->chunk, h->n_local - rg.first);
  rg.padded = (int)round_up(rg.count, 128);
  return rg;
}

static int validate_io(marf_handle* h, const marf_step_io* io) {
  if (!h) return MARF_ERR_INVALID;
  if (!io) return fail(h, MARF_ERR_INVALID, "null io");
  if (!io->mlp_w || !io->mlp_b || !io->warp || !io->rgb || !io->loss_sums) return fail(h, MARF_ERR_INVALID, "missing required pointer");
  if (h->cfg.mask_mode == MARF_MASK_DISK && !io->masks) return fail(h, MARF_ERR_INVALID, "mask_mode=disk needs io->masks");
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT && (!io->mask_w || !io->mask_b || !io->embed))
    return fail(h, MARF_ERR_INVALID, "mask_mode=implicit needs mask_w/mask_b/embed");
  if (h->cfg.use_edges && !io->edges) return fail(h, MARF_ERR_INVALID, "use_edges needs io->edges");
  if (h->cfg.use_edges && h->cfg.mask_mode == MARF_MASK_DISK && !io->masks_eroded)
    return fail(h, MARF_ERR_INVALID, "use_edges with disk masks needs io->masks_eroded");
  return MARF_OK;
}

// data-dependent caches: static mask sums, mask-head input features
static int refresh_data(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  if (io->data_version == h->data_version_seen) return MARF_OK;
  h->data_version_seen = io->data_version;
  h->feats_valid = false;
  CUDA_TRY(h, cudaMemsetAsync(h->sums_static, 0, 4 * sizeof(double), st));
  CUDA_TRY(h, cudaMemsetAsync(h->bad_index, 0, sizeof(double), st));
  if (h->cfg.mask_mode == MARF_MASK_DISK) {
    launch_k(k_sum_f32, 296, 256, 0, st, io->masks, h->n_local, 3.0, h->sums_static + 0);
    LAUNCH_CHECK(h);
    if (io->masks_eroded) {
      launch_k(k_sum_f32, 296, 256, 0, st, io->masks_eroded, h->n_local, 3.0, h->sums_static + 1);
      LAUNCH_CHECK(h);
    }
  }
  return MARF_OK;
}

static int forward_chunk(marf_handle* h, const marf_step_io* io, cudaStream_t st, int ci, bool stats) {
  PxRange rg = chunk_range(h, ci);
  launch_k(k_encode, (rg.padded + 127) / 128, 128, 0, st, h->geo, rg, h->Hm, 0, h->img.act[0], h->img.ld_in[0]);
  LAUNCH_CHECK(h);
  int rc = chain_forward(h, st, h->img, rg.padded);
  if (rc) return rc;
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT) {
    if (!(h->feats_valid && h->n_chunks == 1)) {
      launch_k(k_mask_features, std::min((rg.padded + 7) / 8, 16 * h->n_sms), 256, 0, st, h->geo, rg, io->rgb, io->embed, h->cfg.mask_embed_dim, h->cfg.mask_n_vocab,
               h->cfg.mask_uv_freqs, h->msk.act[0], h->msk.ld_in[0], h->bad_index);
      LAUNCH_CHECK(h);
      h->feats_valid = h->n_chunks == 1;
    }
    rc = chain_forward(h, st, h->msk, rg.padded);
    if (rc) return rc;
  }
  if (stats) {
    LossArgs a;
    a.mask_mode = h->cfg.mask_mode;
    a.logits = h->img.act[h->img.n]; a.ld = h->img.ld_out[h->img.n - 1];
    a.mlogits = h->cfg.mask_mode == MARF_MASK_IMPLICIT ? h->msk.act[h->msk.n] : nullptr;
    a.mld = h->cfg.mask_mode == MARF_MASK_IMPLICIT ? h->msk.ld_out[h->msk.n - 1] : 0;
    a.rgb = io->rgb; a.masks = io->masks;
    a.rgb_pred = io->rgb_pred ? io->rgb_pred : h->pred_rgb;
    a.mask_pred = io->mask_pred ? io->mask_pred : h->pred_mask;
    a.bad_index = h->cfg.mask_mode == MARF_MASK_IMPLICIT ? h->bad_index : nullptr;
    launch_k(k_loss_stats, std::min((rg.padded + 255) / 256, 592), 256, 0, st, h->geo, rg, a, io->loss_sums);
    LAUNCH_CHECK(h);
  }
  return MARF_OK;
}

static int edge_pass(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  const marf_config& c = h->cfg;
  const float* pred = io->rgb_pred ? io->rgb_pred : h->pred_rgb;
  double* ep = io->edge_pred ? io->edge_pred : h->edge_pred;
  EdgeArgs e;
  e.mask_mode = c.mask_mode;
  e.edge_pred = ep; e.edge_label = io->edges; e.label_channels = c.edge_label_channels > 0 ? c.edge_label_channels : 1;
  e.masks_eroded = io->masks_eroded;
  e.mask_pred = io->mask_pred ? io->mask_pred : h->pred_mask;
  if (c.rows >= 8 && h->w >= 8 && c.batch * 3 <= 65535 && !getenv("MARF_EDGE_SPLIT")) {
    // Sobel -> Gauss -> statistics in one launch (the reflected window positions stay inside a block's window for images >= 8x8)
    const dim3 ef((unsigned)((h->w + kEfW - 1) / kEfW), (unsigned)((c.rows + kEfH - 1) / kEfH), (unsigned)(c.batch * 3));
    launch_k(k_edge_fused, ef, 256, 0, st, pred, c.rows, h->w, e, ep, io->loss_sums);
    LAUNCH_CHECK(h);
    return MARF_OK;
  }
  const dim3 eg((unsigned)((c.rows * h->w + 255) / 256), (unsigned)(c.batch * 3));     // x: one plane, y: image * 3 + channel
  launch_k(k_sobel_mag, eg, 256, 0, st, pred, c.batch, 3, c.rows, h->w, 1, h->edge_mag);
  LAUNCH_CHECK(h);
  launch_k(k_gauss5, eg, 256, 0, st, h->edge_mag, c.batch * 3, c.rows, h->w, ep);
  LAUNCH_CHECK(h);
  launch_k(k_edge_stats, (unsigned)std::min<long long>((h->n_local + 255) / 256, 592), 256, 0, st, h->geo, h->n_local, e, io->loss_sums);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

static int backward_chunk(marf_handle* h, const marf_step_io* io, cudaStream_t st, int ci) {
  PxRange rg = chunk_range(h, ci);
  const marf_config& c = h->cfg;
  bool implicit = c.mask_mode == MARF_MASK_IMPLICIT;
  GradArgs ga;
  ga.l.mask_mode = c.mask_mode;
  ga.l.logits = h->img.act[h->img.n]; ga.l.ld = h->img.ld_out[h->img.n - 1];
  ga.l.mlogits = implicit ? h->msk.act[h->msk.n] : nullptr;
  ga.l.mld = implicit ? h->msk.ld_out[h->msk.n - 1] : 0;
  ga.l.rgb = io->rgb; ga.l.masks = io->masks; ga.l.rgb_pred = nullptr; ga.l.mask_pred = nullptr; ga.l.bad_index = nullptr;
  ga.c_rgb = io->c_rgb; ga.c_mask = io->c_mask; ga.c_edge = io->c_edge;
  ga.edge_pred = (implicit && c.use_edges) ? (io->edge_pred ? io->edge_pred : h->edge_pred) : nullptr;
  ga.edge_label = io->edges; ga.label_channels = c.edge_label_channels > 0 ? c.edge_label_channels : 1;
  ga.dlogits = h->dYa; ga.dld = h->img.ld_out[h->img.n - 1];
  // the mask head's last-layer gradient is staged in its own logits buffer's twin: reuse dYb tail is unsafe,
  // so it is written after the image chain has consumed dYa (two launches of the same kernel).
  ga.dmlogits = nullptr; ga.dmld = 0;
  ga.dl_bf16 = nullptr; ga.dml_bf16 = nullptr;
  ga.sums = nullptr; ga.norm_rgb = 0; ga.norm_edge = 0; ga.use_edges = 0;      // (fp32 path: coefficients from k_loss_coef)
  launch_k(k_loss_grad, (rg.padded + 127) / 128, 128, 0, st, h->geo, rg, ga, h->coef);
  LAUNCH_CHECK(h);
  float* dx0 = nullptr;
  int rc = chain_backward(h, st, h->img, rg.padded, &dx0);
  if (rc) return rc;
  launch_k(k_encode_backward, (rg.padded + 255) / 256, 256, 0, st, h->geo, rg, h->Hm, dx0, h->img.ld_in[0], h->G);
  LAUNCH_CHECK(h);
  if (implicit) {
    ga.dlogits = h->dYb;                       // scratch (ignored)
    ga.dmlogits = h->dYa; ga.dmld = h->msk.ld_out[h->msk.n - 1];
    launch_k(k_loss_grad, (rg.padded + 127) / 128, 128, 0, st, h->geo, rg, ga, h->coef);
    LAUNCH_CHECK(h);
    rc = chain_backward(h, st, h->msk, rg.padded, nullptr);
    if (rc) return rc;
  }
  return MARF_OK;
}

static int begin_step(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool pack_fp32 = true, bool sl3 = true) {
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  int rc = refresh_data(h, io, st);
  if (rc) return rc;
  set_schedule(h, io->progress);
  if (pack_fp32) {
    rc = pack_chain(h, st, h->img, io->mlp_w, io->mlp_b);
    if (rc) return rc;
    if (h->cfg.mask_mode == MARF_MASK_IMPLICIT) {
      rc = pack_chain(h, st, h->msk, io->mask_w, io->mask_b);
      if (rc) return rc;
    }
  }
  if (sl3) {
    launch_k(k_sl3_to_SL3, h->cfg.batch_global, 64, 0, st, io->warp, h->cfg.batch_global, h->Hm);
    LAUNCH_CHECK(h);
  }
  return MARF_OK;
}

static int begin_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st, const double* n_rgb_dev) {
  CUDA_TRY(h, cudaMemsetAsync(h->img.gWp[0], 0, chain_param_floats(h->img) * sizeof(float), st));
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT)
    CUDA_TRY(h, cudaMemsetAsync(h->msk.gWp[0], 0, chain_param_floats(h->msk) * sizeof(float), st));
  CUDA_TRY(h, cudaMemsetAsync(h->G, 0, (size_t)h->cfg.batch * 9 * sizeof(double), st));
  (void)n_rgb_dev;
  return MARF_OK;
}

static int finish_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool unpack = true) {
  if (!io->g_warp) return fail(h, MARF_ERR_INVALID, "missing g_warp");
  CUDA_TRY(h, cudaMemsetAsync(io->g_warp, 0, (size_t)h->cfg.batch_global * 8 * sizeof(float), st));
  launch_k(k_sl3_backward, h->cfg.batch, 64, 0, st, io->warp, h->G, h->cfg.patch_offset, h->cfg.batch, io->g_warp);
  LAUNCH_CHECK(h);
  if (!unpack) return MARF_OK;
  int rc = unpack_chain(h, st, h->img, io->g_mlp_w, io->g_mlp_b);
  if (rc) return rc;
  if (h->cfg.mask_mode == MARF_MASK_IMPLICIT) rc = unpack_chain(h, st, h->msk, io->g_mask_w, io->g_mask_b);
  return rc;
}

namespace marf {
int engine_begin_step(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool pack_fp32, bool sl3) { return begin_step(h, io, st, pack_fp32, sl3); }
int engine_edge_pass(marf_handle* h, const marf_step_io* io, cudaStream_t st) { return edge_pass(h, io, st); }
int engine_begin_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st) { return begin_backward(h, io, st, nullptr); }
int engine_finish_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st, bool unpack) { return finish_backward(h, io, st, unpack); }
}  // namespace marf

static int fp32_forward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  int rc = begin_step(h, io, st);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemsetAsync(io->loss_sums, 0, MARF_N_SUMS * sizeof(double), st));
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    rc = forward_chunk(h, io, st, ci, true);
    if (rc) return rc;
  }
  if (h->cfg.use_edges) {
    rc = edge_pass(h, io, st);
    if (rc) return rc;
  }
  h->acts_valid = h->n_chunks == 1;
  return MARF_OK;
}

static int fp32_backward(marf_handle* h, const marf_step_io* io, cudaStream_t st) {
  int rc = begin_backward(h, io, st, nullptr);
  if (rc) return rc;
  launch_k(k_loss_coef, 1, 1, 0, st, io->loss_sums, io->norm_rgb, io->norm_edge, h->cfg.use_edges, h->coef);
  LAUNCH_CHECK(h);
  for (int ci = 0; ci < h->n_chunks; ++ci) {
    if (!h->acts_valid) {
      rc = forward_chunk(h, io, st, ci, false);
      if (rc) return rc;
    }
    rc = backward_chunk(h, io, st, ci);
    if (rc) return rc;
  }
  h->acts_valid = false;
  return finish_backward(h, io, st);
}

extern "C" int marf_step_forward(marf_handle* h, const marf_step_io* io, void* stream) {
  int rc = validate_io(h, io);
  if (rc) return rc;
  if (h->cfg.precision == MARF_BF16) return bf16_forward(h, io, (cudaStream_t)stream);
  return fp32_forward(h, io, (cudaStream_t)stream);
}

extern "C" int marf_step_backward(marf_handle* h, const marf_step_io* io, void* stream) {
  int rc = validate_io(h, io);
  if (rc) return rc;
  if (h->cfg.precision == MARF_BF16) return bf16_backward(h, io, (cudaStream_t)stream);
  return fp32_backward(h, io, (cudaStream_t)stream);
}

extern "C" int marf_step(marf_handle* h, const marf_step_io* io, void* stream) {
  int rc = validate_io(h, io);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->cfg.precision == MARF_BF16) {
    std::string why;
    if (!bf16_supported(h, io, &why)) return fail(h, MARF_ERR_UNSUPPORTED, "precision=bf16: " + why);
    return bf16_step(h, io, st);
  }
  rc = fp32_forward(h, io, st);
  if (rc) return rc;
  return fp32_backward(h, io, st);
}

// ------------------------------------------------------------------------------------------------
extern "C" int marf_render(marf_handle* h, const marf_render_io* io, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!io || !io->mlp_w || !io->mlp_b || !io->rgb || io->n_patches <= 0) return fail(h, MARF_ERR_INVALID, "bad render io");
  if (io->warp && io->n_patches > h->cfg.batch_global) return fail(h, MARF_ERR_INVALID, "render: n_patches > batch_global");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  set_schedule(h, io->progress);
  int rc = pack_chain(h, st, h->img, io->mlp_w, io->mlp_b);
  if (rc) return rc;
  Geo g = h->geo;
  const marf_config& c = h->cfg;
  g.h = io->crop ? c.patch_H : c.H;
  g.w = io->crop ? c.patch_W : c.W;
  g.y0 = io->crop ? c.H / 2 - c.patch_H / 2 : 0;
  g.x0 = io->crop ? c.W / 2 - c.patch_W / 2 : 0;
  g.rows = g.h; g.row_offset = 0; g.patch_offset = 0;
  if (io->warp) {
    launch_k(k_sl3_to_SL3, io->n_patches, 64, 0, st, io->warp, io->n_patches, h->Hm);
    LAUNCH_CHECK(h);
  }
  long long n = (long long)io->n_patches * g.h * g.w;
  for (long long first = 0; first < n; first += h->render_rows) {
    PxRange rg;
    rg.first = first;
    rg.count = (int)std::min<long long>(h->render_rows, n - first);
    rg.padded = (int)round_up(rg.count, 128);
    launch_k(k_encode, (rg.padded + 127) / 128, 128, 0, st, g, rg, h->Hm, io->warp ? 0 : 1, h->img.act[0], h->img.ld_in[0]);
    LAUNCH_CHECK(h);
    rc = chain_forward(h, st, h->img, rg.padded);
    if (rc) return rc;
    launch_k(k_sigmoid_out, (rg.count + 255) / 256, 256, 0, st, rg.count, h->img.act[h->img.n], h->img.ld_out[h->img.n - 1],
                                                          io->rgb + first * 3);
    LAUNCH_CHECK(h);
  }
  h->acts_valid = false;
  h->feats_valid = h->feats_valid;   // mask features live in the mask chain's buffers: untouched
  return MARF_OK;
}

// NeuralImageFunction.forward(coord_2d) for explicit coordinates (model/planar.py:429-449): fp32 arithmetic in both precision
// modes (like marf_render: the forward-only path is the <=1e-3 output-parity path)
extern "C" int marf_forward_points(marf_handle* h, const float* const* mlp_w, const float* const* mlp_b, const float* xy, int64_t n,
                                   float progress, float* rgb, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!mlp_w || !mlp_b || !xy || !rgb || n <= 0) return fail(h, MARF_ERR_INVALID, "bad forward_points args");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h, cudaSetDevice(h->cfg.device));
  set_schedule(h, progress);
  int rc = pack_chain(h, st, h->img, mlp_w, mlp_b);
  if (rc) return rc;
  for (int64_t first = 0; first < n; first += h->render_rows) {
    const int count = (int)std::min<int64_t>(h->render_rows, n - first);
    const int padded = (int)round_up(count, 128);
    launch_k(k_encode_points, (padded + 127) / 128, 128, 0, st, h->geo, xy + 2 * first, count, padded, h->img.act[0], h->img.ld_in[0]);
    LAUNCH_CHECK(h);
    rc = chain_forward(h, st, h->img, padded);
    if (rc) return rc;
    launch_k(k_sigmoid_out, (count + 255) / 256, 256, 0, st, count, h->img.act[h->img.n], h->img.ld_out[h->img.n - 1], rgb + first * 3);
    LAUNCH_CHECK(h);
  }
  h->acts_valid = false;
  return MARF_OK;
}

extern "C" int marf_sl3_to_SL3(marf_handle* h, const float* warp, int32_t n, float* out9, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!warp || !out9 || n <= 0) return fail(h, MARF_ERR_INVALID, "bad sl3 args");
  launch_k(k_sl3_to_SL3, n, 64, 0, (cudaStream_t)stream, warp, n, out9);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

extern "C" int marf_warp_corners(marf_handle* h, const float* warp, int32_t n, float* out, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!warp || !out || n <= 0 || n > h->cfg.batch_global) return fail(h, MARF_ERR_INVALID, "bad corner args");
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(k_sl3_to_SL3, n, 64, 0, st, warp, n, h->Hm);
  LAUNCH_CHECK(h);
  Geo g = h->geo;
  g.h = h->cfg.patch_H; g.w = h->cfg.patch_W;
  launch_k(k_warp_corners, (n * 4 + 63) / 64, 64, 0, st, g, h->Hm, n, out);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

// ------------------------------------------------------------------------------------------------ fused Adam
struct AdamEntry { float* p; const float* g; float* m; float* v; long long n; float lr; int zero_n; };
constexpr int kMaxAdam = 40;
struct AdamTable { AdamEntry e[kMaxAdam]; float beta1, beta2, eps, bc1, bc2_sqrt; };

__global__ void k_adam(const __grid_constant__ AdamTable t) {
  pdl_wait();
  const AdamEntry& E = t.e[blockIdx.y];
  const float step_size = E.lr / t.bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < E.n; i += (long long)gridDim.x * blockDim.x) {
    const float g = E.g[i];
    float m = E.m[i], v = E.v[i];
    m = m + (g - m) * (1.0f - t.beta1);                       // exp_avg.lerp_(grad, 1-beta1)
    v = v * t.beta2 + g * g * (1.0f - t.beta2);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1-beta2)
    const float denom = sqrtf(v) / t.bc2_sqrt + t.eps;
    float p = E.p[i] - step_size * (m / denom);
    if (i < E.zero_n) p = 0.0f;                               // warp.fix_first
    E.m[i] = m; E.v[i] = v; E.p[i] = p;
  }
}

// loss scalars from the sums the step left on the device (model/planar.py:362-378 and :172-185) in one launch:
// out = {rgb, mask, edge, render = (1-alpha) rgb + 0.5 mask + alpha edge, all = sum_k weight_k loss_k}
static __global__ void k_loss_scalars(const double* __restrict__ sums, int implicit, int use_edges, double alpha, double w_render,
                                      double w_rgb, double w_mask, double w_edge, double* __restrict__ out) {
  pdl_wait();
  const double rgb = sums[MARF_S_RGB] / sums[MARF_N_RGB];
  const double mask = implicit ? sums[MARF_S_MASK] / sums[MARF_N_MASK] : 0.0;
  const double edge = use_edges ? sums[MARF_S_EDGE] / sums[MARF_N_EDGE] : 0.0;
  const double render = (1.0 - alpha) * rgb + 0.5 * mask + alpha * edge;
  out[0] = rgb; out[1] = mask; out[2] = edge; out[3] = render;
  out[4] = w_render * render + w_rgb * rgb + w_mask * mask + w_edge * edge;
}

extern "C" int marf_loss_scalars(marf_handle* h, const double* sums, double alpha, const double* weights4, double* out5, void* stream) {
  if (!h || !sums || !weights4 || !out5) return MARF_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  launch_k(k_loss_scalars, 1, 1, 0, st, sums, h->cfg.mask_mode == MARF_MASK_IMPLICIT ? 1 : 0, h->cfg.use_edges ? 1 : 0, alpha,
           weights4[0], weights4[1], weights4[2], weights4[3], out5);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

extern "C" int marf_adam_step(marf_handle* h, const marf_adam_io* io, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!io || io->n_tensors <= 0 || io->n_tensors > kMaxAdam || !io->params || !io->grads || !io->exp_avg || !io->exp_avg_sq ||
      !io->numel || !io->lr || io->step < 1)
    return fail(h, MARF_ERR_INVALID, "bad adam io");
  AdamTable t;
  long long max_n = 0;
  for (int i = 0; i < io->n_tensors; ++i) {
    t.e[i].p = io->params[i]; t.e[i].g = io->grads[i]; t.e[i].m = io->exp_avg[i]; t.e[i].v = io->exp_avg_sq[i];
    t.e[i].n = io->numel[i]; t.e[i].lr = io->lr[i];
    t.e[i].zero_n = (i == io->zero_tensor) ? io->zero_count : 0;
    max_n = std::max<long long>(max_n, io->numel[i]);
  }
  t.beta1 = io->beta1; t.beta2 = io->beta2; t.eps = io->eps;
  t.bc1 = (float)(1.0 - pow((double)io->beta1, (double)io->step));
  t.bc2_sqrt = (float)sqrt(1.0 - pow((double)io->beta2, (double)io->step));
  dim3 grid((unsigned)std::min<long long>((max_n + 255) / 256, 64), io->n_tensors);
  launch_k(k_adam, grid, 256, 0, (cudaStream_t)stream, t);
  LAUNCH_CHECK(h);
  return MARF_OK;
}

__global__ void k_warp_points(const float* __restrict__ xy, const float* __restrict__ Hm, int n, int p,
                              float* __restrict__ out) {
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * p) return;
  int b = (int)(i / p);
  float u, v, qz;
  apply_h(Hm + 9 * b, xy[2 * i], xy[2 * i + 1], u, v, qz);
  out[2 * i] = u;
  out[2 * i + 1] = v;
}

extern "C" int marf_warp_points(marf_handle* h, const float* xy, const float* warp, int32_t n, int32_t p, float* out,
                                void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!xy || !warp || !out || n <= 0 || p <= 0) return fail(h, MARF_ERR_INVALID, "bad warp_points args");
  cudaStream_t st = (cudaStream_t)stream;
  float* Hm = nullptr;
  CUDA_TRY(h, cudaMallocAsync((void**)&Hm, (size_t)n * 9 * sizeof(float), st));
  launch_k(k_sl3_to_SL3, n, 64, 0, st, warp, n, Hm);
  LAUNCH_CHECK(h);
  long long tot = (long long)n * p;
  launch_k(k_warp_points, (unsigned)((tot + 255) / 256), 256, 0, st, xy, Hm, n, p, out);
  LAUNCH_CHECK(h);
  CUDA_TRY(h, cudaFreeAsync(Hm, st));
  return MARF_OK;
}

extern "C" int marf_compute_edges(marf_handle* h, const float* images, int32_t n, int32_t c, int32_t rows, int32_t w,
                                  double* out, void* stream) {
  if (!h) return MARF_ERR_INVALID;
  if (!images || !out || n <= 0 || c <= 0 || rows <= 0 || w <= 0) return fail(h, MARF_ERR_INVALID, "bad edge args");
  cudaStream_t st = (cudaStream_t)stream;
  long long tot = (long long)n * c * rows * w;
  if ((long long)n * c > 65535) return fail(h, MARF_ERR_INVALID, "marf_compute_edges: more than 65535 planes");
  double* mag = nullptr;
  CUDA_TRY(h, cudaMallocAsync((void**)&mag, tot * sizeof(double), st));
  const dim3 eg((unsigned)((rows * w + 255) / 256), (unsigned)(n * c));
  launch_k(k_sobel_mag, eg, 256, 0, st, images, n, c, rows, w, 0, mag);
  LAUNCH_CHECK(h);
  launch_k(k_gauss5, eg, 256, 0, st, mag, n * c, rows, w, out);
  LAUNCH_CHECK(h);
  CUDA_TRY(h, cudaFreeAsync(mag, st));
  return MARF_OK;
}
