// fp32 (CUDA-core) kernels of the planar step: geometry, encodings, SGEMM family, losses.
// This is the parity mode (MARF_FP32): every arithmetic step is fp32 like the reference's
// eager PyTorch path, reductions are fp64.
#pragma once
#include "common.cuh"

namespace marf {

// ============================================================================================
// sl(3) -> SL(3): H = expm(A(h))  (warp.py:98-106), and its adjoint for the backward pass
// (SURVEY.md §8 a-3): dA = expm([[A^T, G],[0, A^T]])[:3,3:].  One thread per patch, fp64 inside.
// ============================================================================================
template <int N>
__device__ void expm_small(const double* A, double* E) {
  double nrm = 0.0;
  for (int j = 0; j < N; ++j) {
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += fabs(A[i * N + j]);
    nrm = fmax(nrm, s);
  }
  int sq = 0;
  while (nrm > 0.25 && sq < 60) { nrm *= 0.5; ++sq; }
  double scale = ldexp(1.0, -sq);
  double X[N * N], T[N * N], R[N * N];
  for (int i = 0; i < N * N; ++i) { X[i] = A[i] * scale; T[i] = (i / N == i % N) ? 1.0 : 0.0; R[i] = T[i]; }
  for (int k = 1; k <= 16; ++k) {            // Taylor: 0.25^17/17! < 1e-24
    double Tn[N * N];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int q = 0; q < N; ++q) s += T[i * N + q] * X[q * N + j];
        Tn[i * N + j] = s / (double)k;
      }
    for (int i = 0; i < N * N; ++i) { T[i] = Tn[i]; R[i] += Tn[i]; }
  }
  for (int s = 0; s < sq; ++s) {
    double Rn[N * N];
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        double acc = 0.0;
        for (int q = 0; q < N; ++q) acc += R[i * N + q] * R[q * N + j];
        Rn[i * N + j] = acc;
      }
    for (int i = 0; i < N * N; ++i) R[i] = Rn[i];
  }
  for (int i = 0; i < N * N; ++i) E[i] = R[i];
}

__device__ __forceinline__ void sl3_generator(const float* h, double* A) {
  // A = [[h5,h3,h1],[h4,-h5-h6,h2],[h7,h8,h6]], h1..h8 = h[0..7]  (warp.py:100-104)
  A[0] = h[4]; A[1] = h[2]; A[2] = h[0];
  A[3] = h[3]; A[4] = -(double)h[4] - (double)h[5]; A[5] = h[1];
  A[6] = h[6]; A[7] = h[7]; A[8] = h[5];
}

// Cooperative expm: N*N threads of one block, one matrix element each, fp64 in shared memory.
// Scaling-and-squaring with a degree-12 Taylor polynomial (||A/2^s||_1 <= 0.25 -> truncation < 1e-16).
template <int N>
__device__ void expm_coop(double* X /*[N*N] in: A, scratch*/, double* T, double* R, int tid) {
  const bool on = tid < N * N;
  const int i = tid / N, j = tid % N;
  __shared__ int s_sq;
  if (tid == 0) {
    double nrm = 0.0;
    for (int c = 0; c < N; ++c) {
      double cs = 0.0;
      for (int r = 0; r < N; ++r) cs += fabs(X[r * N + c]);
      nrm = fmax(nrm, cs);
    }
    int sq = 0;
    while (nrm > 0.25 && sq < 60) { nrm *= 0.5; ++sq; }
    s_sq = sq;
  }
  __syncthreads();
  const int sq = s_sq;
  if (on) {
    X[tid] *= ldexp(1.0, -sq);
    T[tid] = (i == j) ? 1.0 : 0.0;
    R[tid] = T[tid];
  }
  __syncthreads();
  for (int k = 1; k <= 12; ++k) {
    double v = 0.0;
    if (on) {
      for (int q = 0; q < N; ++q) v += T[i * N + q] * X[q * N + j];
      v /= (double)k;
    }
    __syncthreads();
    if (on) { T[tid] = v; R[tid] += v; }
    __syncthreads();
  }
  for (int s = 0; s < sq; ++s) {
    double v = 0.0;
    if (on) for (int q = 0; q < N; ++q) v += R[i * N + q] * R[q * N + j];
    __syncthreads();
    if (on) R[tid] = v;
    __syncthreads();
  }
}

// one block (64 threads) per patch
static __global__ void k_sl3_to_SL3(const float* __restrict__ warp, int n, float* __restrict__ out9) {
  pdl_wait();
  __shared__ double X[9], T[9], R[9];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    if (warp) sl3_generator(warp + 8 * b, X);
    else for (int i = 0; i < 9; ++i) X[i] = 0.0;
  }
  __syncthreads();
  expm_coop<3>(X, T, R, tid);
  if (tid < 9) out9[9 * b + tid] = (float)R[tid];
}

// G[bl] = dL/dH (row-major 3x3, fp64) -> g_warp[bl + patch_offset, 8] (fp32); all threads of the block call it (>= 36 threads)
static __device__ void sl3_backward_block(const float* __restrict__ warp, const double* __restrict__ G, int patch_offset, int bl,
                                   float* __restrict__ g_warp) {
  __shared__ double X[36], T[36], R[36];
  const int tid = threadIdx.x;
  const int b = bl + patch_offset;
  __syncthreads();                                   // (the shared scratch may still be read by a previous call)
  if (tid == 0) {
    double A[9];
    sl3_generator(warp + 8 * b, A);
    for (int i = 0; i < 36; ++i) X[i] = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        X[i * 6 + j] = A[j * 3 + i];                 // A^T
        X[(i + 3) * 6 + (j + 3)] = A[j * 3 + i];     // A^T
        X[i * 6 + (j + 3)] = G[9 * bl + i * 3 + j];  // upstream
      }
  }
  __syncthreads();
  expm_coop<6>(X, T, R, tid);
  if (tid == 0) {
    double dA[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) dA[i * 3 + j] = R[i * 6 + (j + 3)];
    float* o = g_warp + 8 * b;
    o[0] = (float)dA[2];            // h1 -> A02
    o[1] = (float)dA[5];            // h2 -> A12
    o[2] = (float)dA[1];            // h3 -> A01
    o[3] = (float)dA[3];            // h4 -> A10
    o[4] = (float)(dA[0] - dA[4]);  // h5 -> A00, -A11
    o[5] = (float)(dA[8] - dA[4]);  // h6 -> A22, -A11
    o[6] = (float)dA[6];            // h7 -> A20
    o[7] = (float)dA[7];            // h8 -> A21
  }
}

// one block per patch
static __global__ void k_sl3_backward(const float* __restrict__ warp, const double* __restrict__ G, int patch_offset,
                               int n_local, float* __restrict__ g_warp) {
  pdl_wait();
  (void)n_local;
  sl3_backward_block(warp, G, patch_offset, blockIdx.x, g_warp);
}

// warped crop corners (warp.py:83-93): [(X0,Y0),(X0,Y1),(X1,Y1),(X1,Y0)]
static __global__ void k_warp_corners(Geo g, const float* __restrict__ Hm, int n, float* __restrict__ out) {
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 4) return;
  int b = i / 4, c = i % 4;
  // python doubles rounded to f32 once (torch.tensor(corners, dtype=float32))
  double Y[2], X[2];
  int yc[2] = {g.H / 2 - g.h / 2, g.H / 2 + g.h / 2};
  int xc[2] = {g.W / 2 - g.w / 2, g.W / 2 + g.w / 2};
  double nh = (double)g.H / (double)max(g.H, g.W), nw = (double)g.W / (double)max(g.H, g.W);
  for (int k = 0; k < 2; ++k) {
    Y[k] = ((yc[k] + 0.5) / g.H * 2 - 1) * nh;
    X[k] = ((xc[k] + 0.5) / g.W * 2 - 1) * nw;
  }
  const int xi[4] = {0, 0, 1, 1}, yi[4] = {0, 1, 1, 0};
  float x = (float)X[xi[c]], y = (float)Y[yi[c]], u, v, qz;
  apply_h(Hm + 9 * b, x, y, u, v, qz);
  out[2 * i] = u;
  out[2 * i + 1] = v;
}

// ============================================================================================
// encoding prologue: grid -> warp -> [u, v, w_k sin(f_k u), w_k cos(f_k u), w_k sin(f_k v), w_k cos(f_k v)]
// (warp.py:33-81, model/planar.py:451-471,434).  One thread per pixel-sample; X0 is [n_pad, ld].
// ============================================================================================
static __global__ void k_encode(Geo g, PxRange rg, const float* __restrict__ Hm /* [batch_global,9] or identity */,
                         int identity, float* __restrict__ X0, int ld) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rg.padded) return;
  float* o = X0 + (size_t)t * ld;
  if (t >= rg.count) {
    for (int j = 0; j < ld; ++j) o[j] = 0.0f;
    return;
  }
  int b, r, c;
  decode_px(g, rg.first + t, b, r, c);
  float x, y, u, v, qz;
  grid_xy(g, r, c, x, y);
  if (identity) { u = x; v = y; }
  else apply_h(Hm + 9 * (b + g.patch_offset), x, y, u, v, qz);
  o[0] = u;
  o[1] = v;
  const int L = g.L;
  for (int k = 0; k < L; ++k) {
    float su, cu, sv, cv;
    sincosf(u * g.band_f[k], &su, &cu);
    sincosf(v * g.band_f[k], &sv, &cv);
    float wk = g.band_w[k];
    o[2 + k] = su * wk;
    o[2 + L + k] = cu * wk;
    o[2 + 2 * L + k] = sv * wk;
    o[2 + 3 * L + k] = cv * wk;
  }
  for (int j = g.d_in; j < ld; ++j) o[j] = 0.0f;
}

// the same for caller-supplied (already warped) coordinates: NeuralImageFunction.forward(coord_2d), model/planar.py:429-436
static __global__ void k_encode_points(Geo g, const float* __restrict__ xy, int count, int padded, float* __restrict__ X0, int ld) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= padded) return;
  float* o = X0 + (size_t)t * ld;
  if (t >= count) {
    for (int j = 0; j < ld; ++j) o[j] = 0.0f;
    return;
  }
  const float u = xy[2 * (size_t)t], v = xy[2 * (size_t)t + 1];
  o[0] = u;
  o[1] = v;
  const int L = g.L;
  for (int k = 0; k < L; ++k) {
    float su, cu, sv, cv;
    sincosf(u * g.band_f[k], &su, &cu);
    sincosf(v * g.band_f[k], &sv, &cv);
    float wk = g.band_w[k];
    o[2 + k] = su * wk;
    o[2 + L + k] = cu * wk;
    o[2 + 2 * L + k] = sv * wk;
    o[2 + 3 * L + k] = cv * wk;
  }
  for (int j = g.d_in; j < ld; ++j) o[j] = 0.0f;
}

// backward of the prologue: dX0 [n,ld] -> per-patch G = sum_p dq (x) [x,y,1]  (SURVEY.md §8 a-4,a-5)
static __global__ void k_encode_backward(Geo g, PxRange rg, const float* __restrict__ Hm, const float* __restrict__ dX0,
                                  int ld, double* __restrict__ G /* [batch,9] local patches */) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool valid = t < rg.count;
  int b = -1, r = 0, c = 0;
  float acc[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) acc[i] = 0.0f;
  if (valid) {
    decode_px(g, rg.first + t, b, r, c);
    float x, y, u, v, qz;
    grid_xy(g, r, c, x, y);
    apply_h(Hm + 9 * (b + g.patch_offset), x, y, u, v, qz);
    const float* d = dX0 + (size_t)t * ld;
    float gu = d[0], gv = d[1];
    const int L = g.L;
    for (int k = 0; k < L; ++k) {
      float su, cu, sv, cv;
      float f = g.band_f[k];
      sincosf(u * f, &su, &cu);
      sincosf(v * f, &sv, &cv);
      float wf = g.band_w[k] * f;
      gu += wf * (d[2 + k] * cu - d[2 + L + k] * su);
      gv += wf * (d[2 + 2 * L + k] * cv - d[2 + 3 * L + k] * sv);
    }
    float inv = 1.0f / qz;
    float dq0 = gu * inv, dq1 = gv * inv, dq2 = -(gu * u + gv * v) * inv;
    acc[0] = dq0 * x; acc[1] = dq0 * y; acc[2] = dq0;
    acc[3] = dq1 * x; acc[4] = dq1 * y; acc[5] = dq1;
    acc[6] = dq2 * x; acc[7] = dq2 * y; acc[8] = dq2;
  }
  // reduce: warp shuffles -> per-warp slots in shared memory -> one fp64 atomic per (patch, entry) per block
  __shared__ float slot[8][9];
  __shared__ int slot_b[8];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int b0 = __shfl_sync(0xffffffffu, b, 0);   // lane 0 is valid whenever any lane of the warp is (t ascending)
  bool uniform = __all_sync(0xffffffffu, b == b0 || b < 0);
  if (uniform) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float sacc = warp_sum(acc[i]);
      if (lane == 0) slot[wid][i] = sacc;
    }
    if (lane == 0) slot_b[wid] = b0;
  } else {
    if (lane == 0) slot_b[wid] = -1;
    if (valid) {
#pragma unroll
      for (int i = 0; i < 9; ++i) atomicAdd(&G[9 * b + i], (double)acc[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    int cur = -1;
    double tot = 0.0;
    for (int w = 0; w < nw; ++w) {
      int wb = slot_b[w];
      if (wb < 0) continue;
      if (wb != cur) {
        if (cur >= 0) atomicAdd(&G[9 * cur + threadIdx.x], tot);
        cur = wb;
        tot = 0.0;
      }
      tot += (double)slot[w][threadIdx.x];
    }
    if (cur >= 0) atomicAdd(&G[9 * cur + threadIdx.x], tot);
  }
}

// ============================================================================================
// mask-head input features (model/planar.py:342-349, :491-518): [E[trunc r], E[trunc g], E[trunc b], PosEmbedding(xy)]
// one warp per pixel-sample row (8 rows per block, grid-stride), 16-byte copies of the embedding rows; iteration-invariant
// (cached by the engine when possible).
// ============================================================================================
static __global__ void __launch_bounds__(256) k_mask_features(Geo g, PxRange rg, const float* __restrict__ rgb, const float* __restrict__ embed,
                                int embed_dim, int n_vocab, int n_freqs, float* __restrict__ F, int ld, double* __restrict__ bad) {
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long per = (long long)g.rows * g.w;
  const int k_col = 3 * embed_dim, k_uv = 2 + 4 * n_freqs;
  const bool vec = (embed_dim & 3) == 0 && (ld & 3) == 0;
  for (int t = blockIdx.x * 8 + warp; t < rg.padded; t += gridDim.x * 8) {
    float* o = F + (size_t)t * ld;
    if (t >= rg.count) {
      for (int j = lane; j < ld; j += 32) o[j] = 0.0f;
      continue;
    }
    int b, r, c;
    const long long i = rg.first + t;
    decode_px(g, i, b, r, c);
    const long long rem = i - (long long)b * per;
    for (int ch = 0; ch < 3; ++ch) {
      const float val = rgb[((long long)b * 3 + ch) * per + rem];
      long long idx = (long long)val;                 // .long(): truncation toward zero
      if (idx < 0 || idx >= n_vocab) {                // the reference raises IndexError here (nn.Embedding, model/planar.py:344)
        if (lane == 0) atomicAdd(bad, 1.0);
        idx = idx < 0 ? 0 : n_vocab - 1;
      }
      const float* e = embed + idx * embed_dim;
      float* oc = o + ch * embed_dim;
      if (vec) {
        for (int j = lane * 4; j < embed_dim; j += 128) *reinterpret_cast<float4*>(oc + j) = *reinterpret_cast<const float4*>(e + j);
      } else {
        for (int j = lane; j < embed_dim; j += 32) oc[j] = e[j];
      }
    }
    float x, y;
    grid_xy(g, r, c, x, y);
    for (int j = lane; j < k_uv; j += 32) {
      float val;
      if (j < 2) val = j == 0 ? x : y;
      else {
        const int q = j - 2, fi = q >> 2, w4 = q & 3;  // per frequency: sin x, sin y, cos x, cos y
        const float f = (float)(1 << fi);
        const float a = f * ((w4 & 1) ? y : x);
        val = (w4 < 2) ? sinf(a) : cosf(a);
      }
      o[k_col + j] = val;
    }
    for (int j = k_col + k_uv + lane; j < ld; j += 32) o[j] = 0.0f;
  }
}

// ============================================================================================
// SGEMM family.  C[M,N] (+)= A[M,K] * B[K,N] with fp32 FMA, 128 x (16*TN) x 16 tiles, 256 threads.
//   A_KC: A(m,k) = A[m*lda + k]  (K contiguous)      else A(m,k) = A[k*lda + m]  (M contiguous)
//   B_KC: B(k,n) = B[n*ldb + k]  (K contiguous)      else B(k,n) = B[k*ldb + n]  (N contiguous)
// All of M,N,K,lda,ldb,ldc are multiples of 4 (engine pads), pointers 16-byte aligned.
// ============================================================================================
enum Epi { EPI_BIAS = 0, EPI_BIAS_RELU = 1, EPI_PLAIN = 2, EPI_RELU_MASK = 3, EPI_ATOMIC = 4, EPI_ATOMIC_T = 5 };

constexpr int GBM = 128, GBK = 16;

__device__ __forceinline__ int swz(int k, int idx) { return idx ^ (((k >> 2) & 3) << 3); }

template <bool KC, int ROWS>
__device__ __forceinline__ void load_tile(float (*S)[GBM], const float* __restrict__ P, int ld, int row0, int nrows,
                                          int k0, int kend, int tid) {
  // fills S[k][row] for row in [0,ROWS), k in [0,GBK); zero outside [row0+.. < nrows) x [k0+.. < kend)
  if (KC) {
    // float4 along K: ROWS*4 float4 in the tile
    for (int idx = tid; idx < ROWS * (GBK / 4); idx += 256) {
      int row = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int gr = row0 + row, gk = k0 + kq;
      if (gr < nrows && gk < kend) v = *reinterpret_cast<const float4*>(P + (size_t)gr * ld + gk);
      S[kq + 0][swz(kq + 0, row)] = v.x;
      S[kq + 1][swz(kq + 1, row)] = v.y;
      S[kq + 2][swz(kq + 2, row)] = v.z;
      S[kq + 3][swz(kq + 3, row)] = v.w;
    }
  } else {
    // float4 along the row index: GBK * ROWS/4 float4
    for (int idx = tid; idx < GBK * (ROWS / 4); idx += 256) {
      int k = idx / (ROWS / 4), r4 = (idx - k * (ROWS / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int gr = row0 + r4, gk = k0 + k;
      if (gr < nrows && gk < kend) v = *reinterpret_cast<const float4*>(P + (size_t)gk * ld + gr);
      *reinterpret_cast<float4*>(&S[k][swz(k, r4)]) = v;
    }
  }
}

template <bool A_KC, bool B_KC, int TN, int EPI>
static __global__ void __launch_bounds__(256) k_sgemm(int M, int N, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ B, int ldb, float* __restrict__ C, int ldc,
                                               const float* __restrict__ aux, int ldaux, int k_split) {
  constexpr int BN = 16 * TN;
  pdl_wait();
  __shared__ __align__(16) float As[2][GBK][GBM];
  __shared__ __align__(16) float Bs[2][GBK][GBM];   // only the first BN columns are used
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * BN;
  // split-K range (multiples of GBK)
  int kt_total = (K + GBK - 1) / GBK;
  int kt_per = (kt_total + k_split - 1) / k_split;
  int kt_begin = blockIdx.z * kt_per;
  int kt_end = min(kt_total, kt_begin + kt_per);
  if (kt_begin >= kt_end) return;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_tile<A_KC, GBM>(As[0], A, lda, m0, M, kt_begin * GBK, K, tid);
  load_tile<B_KC, BN>(Bs[0], B, ldb, n0, N, kt_begin * GBK, K, tid);
  __syncthreads();
  int buf = 0;
  for (int kt = kt_begin; kt < kt_end; ++kt) {
    if (kt + 1 < kt_end) {
      load_tile<A_KC, GBM>(As[buf ^ 1], A, lda, m0, M, (kt + 1) * GBK, K, tid);
      load_tile<B_KC, BN>(Bs[buf ^ 1], B, ldb, n0, N, (kt + 1) * GBK, K, tid);
    }
#pragma unroll
    for (int k = 0; k < GBK; ++k) {
      float a[8], b[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][swz(k, ty * 4)]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][swz(k, 64 + ty * 4)]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TN == 8) {
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][swz(k, tx * 4)]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][swz(k, 64 + tx * 4)]);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4 % TN] = b1.x; b[5 % TN] = b1.y; b[6 % TN] = b1.z; b[7 % TN] = b1.w;
      } else {
        b[0] = Bs[buf][k][swz(k, tx)];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + (TN == 8 ? (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) : tx);
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == EPI_BIAS) C[(size_t)m * ldc + n] = v + aux[n];
      else if (EPI == EPI_BIAS_RELU) C[(size_t)m * ldc + n] = fmaxf(v + aux[n], 0.f);
      else if (EPI == EPI_PLAIN) C[(size_t)m * ldc + n] = v;
      else if (EPI == EPI_RELU_MASK) C[(size_t)m * ldc + n] = aux[(size_t)m * ldaux + n] > 0.f ? v : 0.f;
      else if (EPI == EPI_ATOMIC) atomicAdd(&C[(size_t)m * ldc + n], v);
      else atomicAdd(&C[(size_t)n * ldc + m], v);   // transposed accumulate
    }
  }
}

// db[j] += sum_m dY[m,j]
static __global__ void k_colsum(int M, int N, const float* __restrict__ dY, int ld, float* __restrict__ db, int rows_per_block) {
  pdl_wait();
  __shared__ float red[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int n = blockIdx.x * 32 + tx;
  int mb = blockIdx.y * rows_per_block, me = min(M, mb + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int m = mb + ty; m < me; m += 8) s += dY[(size_t)m * ld + n];
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][tx];
    atomicAdd(&db[n], tot);
  }
}

// Whole-chain weight pack / gradient unpack in ONE launch each (blockIdx.y = layer): torch layouts <-> the zero-padded
// workspace, plus (3xTF32 path) the big / small tf32 planes of W and of W^T that the tensor-core layers read by TMA.
struct ChainPackJobs {
  const float* W[MARF_MAX_LAYERS];     // [k_out, k_in] torch layout
  const float* b[MARF_MAX_LAYERS];
  float* Wp[MARF_MAX_LAYERS];          // [ld_out, ld_in]
  float* bp[MARF_MAX_LAYERS];          // [ld_out]
  float* sp_f[MARF_MAX_LAYERS];        // [2 * pad16(ld_out), pad32(ld_in)]  (nullptr: layer stays on the CUDA cores)
  float* sp_t[MARF_MAX_LAYERS];        // [2 * pad16(ld_in), pad32(ld_out)]
  int k_out[MARF_MAX_LAYERS], k_in[MARF_MAX_LAYERS], ld_out[MARF_MAX_LAYERS], ld_in[MARF_MAX_LAYERS];
};
__device__ __forceinline__ float tf32_round_bits(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
static __global__ void k_pack_chain(const __grid_constant__ ChainPackJobs J) {
  pdl_wait();
  const int l = blockIdx.y;
  const float* __restrict__ W = J.W[l];
  const int ko = J.k_out[l], ki = J.k_in[l], lo = J.ld_out[l], li = J.ld_in[l];
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < lo * li; i += stride) {
    const int r = i / li, c = i - r * li;
    J.Wp[l][i] = (r < ko && c < ki) ? W[(size_t)r * ki + c] : 0.f;
  }
  for (int i = t0; i < lo; i += stride) J.bp[l][i] = i < ko ? J.b[l][i] : 0.f;
  if (J.sp_f[l]) {
    {   // forward operand: B[n = out][k = in]
      const int nr = (lo + 15) / 16 * 16, kp = (li + 31) / 32 * 32;
      for (int i = t0; i < nr * kp; i += stride) {
        const int n = i / kp, k = i - n * kp;
        const float x = (n < ko && k < ki) ? W[(size_t)n * ki + k] : 0.f;
        const float big = tf32_round_bits(x);
        J.sp_f[l][i] = big;
        J.sp_f[l][(size_t)nr * kp + i] = tf32_round_bits(x - big);
      }
    }
    {   // dX operand: B[n = in][k = out] = W[k][n]; consecutive threads take consecutive n (contiguous in W)
      const int nr = (li + 15) / 16 * 16, kp = (lo + 31) / 32 * 32;
      for (int i = t0; i < nr * kp; i += stride) {
        const int k = i / nr, n = i - k * nr;
        const float x = (n < ki && k < ko) ? W[(size_t)k * ki + n] : 0.f;
        const float big = tf32_round_bits(x);
        J.sp_t[l][(size_t)n * kp + k] = big;
        J.sp_t[l][(size_t)(nr + n) * kp + k] = tf32_round_bits(x - big);
      }
    }
  }
}
struct ChainUnpackJobs {
  const float* gWp[MARF_MAX_LAYERS];
  const float* gbp[MARF_MAX_LAYERS];
  float* gW[MARF_MAX_LAYERS];
  float* gb[MARF_MAX_LAYERS];
  int k_out[MARF_MAX_LAYERS], k_in[MARF_MAX_LAYERS], ld_in[MARF_MAX_LAYERS];
};
static __global__ void k_unpack_chain(const __grid_constant__ ChainUnpackJobs J) {
  pdl_wait();
  const int l = blockIdx.y;
  const int ko = J.k_out[l], ki = J.k_in[l], li = J.ld_in[l];
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < ko * ki; i += stride) {
    const int r = i / ki, c = i - r * ki;
    J.gW[l][i] = J.gWp[l][(size_t)r * li + c];
  }
  for (int i = t0; i < ko; i += stride) J.gb[l][i] = J.gbp[l][i];
}

// copy `cols` columns between row-major buffers (skip-connection concat / split)
static __global__ void k_copy_cols(int M, int cols, const float* __restrict__ src, int lds, int soff, float* __restrict__ dst,
                            int ldd, int doff, int accumulate) {
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * cols) return;
  int m = (int)(i / cols), c = (int)(i - (long long)m * cols);
  float v = src[(size_t)m * lds + soff + c];
  float* d = dst + (size_t)m * ldd + doff + c;
  *d = accumulate ? *d + v : v;
}
// dY_prev[m,j] = dX[m,j] * (act[m,j] > 0) for j < cols
static __global__ void k_relu_mask(int M, int cols, const float* __restrict__ dX, int ldx, const float* __restrict__ act,
                            int lda, float* __restrict__ dY, int ldy) {
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)M * cols) return;
  int m = (int)(i / cols), c = (int)(i - (long long)m * cols);
  dY[(size_t)m * ldy + c] = act[(size_t)m * lda + c] > 0.f ? dX[(size_t)m * ldx + c] : 0.f;
}

// ============================================================================================
// the narrow output layer (3 colour / 1 mask logits) of the fp32 mode: bandwidth-bound row kernels instead of three
// SGEMM-shaped launches (forward N = 3; backward dW [3, K] + db + dX = (dY W) masked by the layer input, one pass over X)
// ============================================================================================
// lane l of a warp holds the 16-byte chunks l, l + 32, ... (KCH of them) of a row of X; W rows live in SMEM
template <int KCH>
static __global__ void __launch_bounds__(256) k_out_forward(int M, int K, int n_out, const float* __restrict__ X, int ldx,
                                                            const float* __restrict__ W, int ldw, const float* __restrict__ b,
                                                            float* __restrict__ Y, int ldy) {
  pdl_wait();
  __shared__ __align__(16) float sW[4][KCH * 128];
  for (int i = threadIdx.x; i < 4 * KCH * 128; i += 256) {
    const int j = i / (KCH * 128), c = i - j * (KCH * 128);
    sW[j][c] = (j < n_out && c < K) ? W[(size_t)j * ldw + c] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * 8;
  for (int m = blockIdx.x * 8 + warp; m < M; m += nwarps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < KCH; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < K) {
        const float4 x = *reinterpret_cast<const float4*>(X + (size_t)m * ldx + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(&sW[j][c]);
          acc[j] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[j]))));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      float4 y;
      y.x = acc[0] + b[0];
      y.y = n_out > 1 ? acc[1] + b[1] : 0.f;
      y.z = n_out > 2 ? acc[2] + b[2] : 0.f;
      y.w = n_out > 3 ? acc[3] + b[3] : 0.f;
      *reinterpret_cast<float4*>(Y + (size_t)m * ldy) = y;
    }
  }
}

// dX[m, c] = X[m, c] > 0 ? sum_j dY[m, j] W[j, c] : 0;  gW[j, c] += sum_m dY[m, j] X[m, c];  gb[j] += sum_m dY[m, j]
template <int KCH>
static __global__ void __launch_bounds__(256) k_out_backward(int M, int K, int n_out, const float* __restrict__ X, int ldx,
                                                             const float* __restrict__ dY, int ldy, const float* __restrict__ W,
                                                             int ldw, float* __restrict__ dX, int lddx, float* __restrict__ gW,
                                                             float* __restrict__ gb) {
  pdl_wait();
  __shared__ __align__(16) float sW[4][KCH * 128];
  __shared__ float sAcc[4][KCH * 128];
  __shared__ float sB[4];
  for (int i = threadIdx.x; i < 4 * KCH * 128; i += 256) {
    const int j = i / (KCH * 128), c = i - j * (KCH * 128);
    sW[j][c] = (j < n_out && c < K) ? W[(size_t)j * ldw + c] : 0.f;
    sAcc[j][c] = 0.f;
  }
  if (threadIdx.x < 4) sB[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = gridDim.x * 8;
  float aw[4][KCH][4];
  float ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < KCH; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) aw[j][i][e] = 0.f;
  // the next row's operands are requested while the current row is processed (one 1 KB row per warp and iteration would
  // leave the loads latency-bound)
  float4 xn[KCH], dn = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_row = [&](int m) {
    dn = *reinterpret_cast<const float4*>(dY + (size_t)m * ldy);      // (ldy = 4: the padded logits row)
#pragma unroll
    for (int i = 0; i < KCH; ++i) {
      const int c = (lane + 32 * i) * 4;
      xn[i] = c < K ? *reinterpret_cast<const float4*>(X + (size_t)m * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int m = blockIdx.x * 8 + warp;
  if (m < M) load_row(m);
  for (; m < M; m += nwarps) {
    const float4 d4 = dn;
    float4 xc[KCH];
#pragma unroll
    for (int i = 0; i < KCH; ++i) xc[i] = xn[i];
    if (m + nwarps < M) load_row(m + nwarps);
    const float d[4] = {d4.x, n_out > 1 ? d4.y : 0.f, n_out > 2 ? d4.z : 0.f, n_out > 3 ? d4.w : 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) ab[j] += d[j];
#pragma unroll
    for (int i = 0; i < KCH; ++i) {
      const int c = (lane + 32 * i) * 4;
      if (c < K) {
        const float xv[4] = {xc[i].x, xc[i].y, xc[i].z, xc[i].w};
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = *reinterpret_cast<const float4*>(&sW[j][c]);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            g[e] = fmaf(d[j], wv[e], g[e]);
            aw[j][i][e] = fmaf(d[j], xv[e], aw[j][i][e]);
          }
        }
        float4 o;
        o.x = xv[0] > 0.f ? g[0] : 0.f; o.y = xv[1] > 0.f ? g[1] : 0.f; o.z = xv[2] > 0.f ? g[2] : 0.f; o.w = xv[3] > 0.f ? g[3] : 0.f;
        *reinterpret_cast<float4*>(dX + (size_t)m * lddx + c) = o;
      }
    }
  }
  // block reduction of the weight-gradient partials, then one atomic per entry and block
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < KCH; ++i) {
      const int c = (lane + 32 * i) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (c < K) atomicAdd(&sAcc[j][c + e], aw[j][i][e]);
    }
  if (lane == 0)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&sB[j], ab[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < n_out * KCH * 128; i += 256) {
    const int j = i / (KCH * 128), c = i - j * (KCH * 128);
    if (c < K) atomicAdd(&gW[(size_t)j * ldw + c], sAcc[j][c]);
  }
  if (threadIdx.x < n_out) atomicAdd(&gb[threadIdx.x], sB[threadIdx.x]);
}

// ============================================================================================
// losses (model/planar.py:355-391) — statistics pass and gradient pass
// ============================================================================================
struct LossArgs {
  int mask_mode;               // marf_mask_mode
  const float* logits; int ld; // image MLP output before sigmoid [n, ld]
  const float* mlogits; int mld; // mask head output before sigmoid [n, mld] (implicit only)
  const float* rgb;            // [batch,3,rows,w]
  const float* masks;          // [batch,1,rows,w] (disk only)
  float* rgb_pred;             // optional [n_total,3]
  float* mask_pred;            // optional [n_total]
  const double* bad_index;     // optional: the handle's sticky bad-colour-index counter, copied to sums[MARF_BAD_INDEX]
};

__device__ __forceinline__ void block_sum_atomic(double v, double* dst, double* red /*[32]*/) {
  // all threads of the block call this; blockDim.x multiple of 32, <= 1024
  v = warp_sum(v);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0 && t != 0.0) atomicAdd(dst, t);
  }
}

static __global__ void k_loss_stats(Geo g, PxRange rg, LossArgs a, double* __restrict__ sums) {
  pdl_wait();
  __shared__ double red[32];
  double s_rgb = 0, n_rgb = 0, s_mask = 0, n_mask = 0, bad = 0;
  const long long per = (long long)g.rows * g.w;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < rg.count; t += gridDim.x * blockDim.x) {
    long long i = rg.first + t;
    long long b = i / per, rem = i - b * per;
    float m = 1.f;
    if (a.mask_mode == MARF_MASK_DISK) m = a.masks[i];
    else if (a.mask_mode == MARF_MASK_IMPLICIT) {
      m = sigmoidf_acc(a.mlogits[(size_t)t * a.mld]);
      if (a.mask_pred) a.mask_pred[i] = m;
      float om = 1.f - m;
      s_mask += (double)(om * om);
    }
    n_mask += 1.0;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float p = sigmoidf_acc(a.logits[(size_t)t * a.ld + c]);
      if (a.rgb_pred) a.rgb_pred[i * 3 + c] = p;
      float d = (p - a.rgb[(b * 3 + c) * per + rem]) * m;
      acc += d * d;
      if (!isfinite(p)) bad += 1.0;
    }
    s_rgb += (double)acc;
    n_rgb += 3.0 * (double)m;
  }
  block_sum_atomic(s_rgb, &sums[MARF_S_RGB], red);
  block_sum_atomic(n_rgb, &sums[MARF_N_RGB], red);
  if (a.mask_mode == MARF_MASK_IMPLICIT) {
    block_sum_atomic(s_mask, &sums[MARF_S_MASK], red);
    block_sum_atomic(n_mask, &sums[MARF_N_MASK], red);
  }
  block_sum_atomic(bad, &sums[MARF_NONFINITE], red);
  if (a.bad_index && rg.first == 0 && blockIdx.x == 0 && threadIdx.x == 0) sums[MARF_BAD_INDEX] = *a.bad_index;
}

// static mask sum (disk masks): N_RGB = 3 * sum m over the local shard
static __global__ void k_sum_f32(const float* __restrict__ x, long long n, double scale, double* __restrict__ out) {
  pdl_wait();
  double s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += (double)x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s * scale);
}

// edge-loss statistics: S_EDGE = sum_c ((e_c - l) m)^2, N_EDGE = 3 sum m   (model/planar.py:366-369)
struct EdgeArgs {
  int mask_mode;
  const double* edge_pred;     // [batch,3,rows,w]
  const double* edge_label;    // [batch,lc,rows,w]
  int label_channels;
  const float* masks_eroded;   // disk mode
  const float* mask_pred;      // implicit mode [n_total]
};
static __global__ void k_edge_stats(Geo g, long long n_total, EdgeArgs a, double* __restrict__ sums) {
  pdl_wait();
  __shared__ double red[32];
  double s = 0, nn = 0;
  const long long per = (long long)g.rows * g.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    long long b = i / per, rem = i - b * per;
    double m = 1.0;
    if (a.mask_mode == MARF_MASK_DISK) m = (double)a.masks_eroded[i];
    else if (a.mask_mode == MARF_MASK_IMPLICIT) m = (double)a.mask_pred[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double l = a.edge_label[(b * a.label_channels + (a.label_channels == 1 ? 0 : c)) * per + rem];
      double d = (a.edge_pred[(b * 3 + c) * per + rem] - l) * m;
      s += d * d;
    }
    nn += 3.0 * m;
  }
  block_sum_atomic(s, &sums[MARF_S_EDGE], red);
  block_sum_atomic(nn, &sums[MARF_N_EDGE], red);
}

// resolve normalisers on device (after an optional all-reduce of `sums`): no host round trip
__device__ __forceinline__ LossCoef make_loss_coef(const double* __restrict__ sums, double norm_rgb_host, double norm_edge_host,
                                                   int use_edges) {
  double n_rgb = norm_rgb_host > 0 ? norm_rgb_host : sums[MARF_N_RGB];
  double n_edge = norm_edge_host > 0 ? norm_edge_host : sums[MARF_N_EDGE];
  LossCoef c;
  c.inv_n_rgb = 1.0 / n_rgb;
  c.s_over_n2 = 3.0 * sums[MARF_S_RGB] / (n_rgb * n_rgb);
  c.inv_n_mask = sums[MARF_N_MASK] > 0 ? 1.0 / sums[MARF_N_MASK] : 0.0;
  c.inv_n_edge = (use_edges && n_edge > 0) ? 1.0 / n_edge : 0.0;
  c.se_over_n2 = (use_edges && n_edge > 0) ? 3.0 * sums[MARF_S_EDGE] / (n_edge * n_edge) : 0.0;
  return c;
}
static __global__ void k_loss_coef(const double* __restrict__ sums, double norm_rgb_host, double norm_edge_host,
                            int use_edges, LossCoef* __restrict__ out) {
  pdl_wait();
  *out = make_loss_coef(sums, norm_rgb_host, norm_edge_host, use_edges);
}

// the same when the rgb normaliser does not depend on the forward pass (no masks / disk masks): from the static mask sum
// (or the host-supplied global one), so that a multi-chunk step can run forward + backward chunk by chunk in one sweep
static __global__ void k_loss_coef_static(const double* __restrict__ sums_static, int mask_mode, double norm_rgb_host,
                                          long long n_local, LossCoef* __restrict__ out) {
  pdl_wait();
  const double n_rgb = norm_rgb_host > 0 ? norm_rgb_host : (mask_mode == MARF_MASK_DISK ? sums_static[0] : 3.0 * (double)n_local);
  LossCoef c;
  c.inv_n_rgb = 1.0 / n_rgb;
  c.s_over_n2 = 0.0; c.inv_n_mask = 0.0; c.inv_n_edge = 0.0; c.se_over_n2 = 0.0;     // (only the mask head's gradient uses them)
  *out = c;
}

struct GradArgs {
  LossArgs l;
  float c_rgb, c_mask, c_edge;
  const double* edge_pred;     // implicit + edges: per-pixel edge residual feeds the mask gradient
  const double* edge_label;
  int label_channels;
  float* dlogits; int dld;     // out [n_pad, dld]
  float* dmlogits; int dmld;   // out [n_pad, dmld] (implicit only)
  __nv_bfloat16* dl_bf16;      // optional bf16 copies [n_pad, 8] for the tensor-core dX / dW of the output layers (the TMA
                               // boxes over them are 64 columns wide: columns >= 8 are out-of-bounds zero fill, not HBM reads)
  __nv_bfloat16* dml_bf16;
  // when the kernel is given no precomputed LossCoef it resolves the coefficients itself from the (all-reduced) sums
  const double* sums; double norm_rgb, norm_edge; int use_edges;
};

static __global__ void k_loss_grad(Geo g, PxRange rg, GradArgs a, const LossCoef* __restrict__ coefp) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rg.padded) return;
  float* dl = a.dlogits + (size_t)t * a.dld;
  float* dm = a.dmlogits ? a.dmlogits + (size_t)t * a.dmld : nullptr;
  uint4* dlb = a.dl_bf16 ? reinterpret_cast<uint4*>(a.dl_bf16 + (size_t)t * 8) : nullptr;
  uint4* dmb = a.dml_bf16 ? reinterpret_cast<uint4*>(a.dml_bf16 + (size_t)t * 8) : nullptr;
  if (t >= rg.count) {
    for (int j = 0; j < a.dld; ++j) dl[j] = 0.f;
    if (dm) for (int j = 0; j < a.dmld; ++j) dm[j] = 0.f;
    if (dlb) dlb[0] = make_uint4(0u, 0u, 0u, 0u);
    if (dmb) dmb[0] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const LossCoef cf = coefp ? *coefp : make_loss_coef(a.sums, a.norm_rgb, a.norm_edge, a.use_edges);
  long long i = rg.first + t;
  long long per = (long long)g.rows * g.w;
  long long b = i / per, rem = i - b * per;
  float m = 1.f;
  if (a.l.mask_mode == MARF_MASK_DISK) m = a.l.masks[i];
  else if (a.l.mask_mode == MARF_MASK_IMPLICIT) m = sigmoidf_acc(a.l.mlogits[(size_t)t * a.l.mld]);
  float k_rgb = (float)(2.0 * (double)a.c_rgb * cf.inv_n_rgb);
  float sum_d2 = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p = sigmoidf_acc(a.l.logits[(size_t)t * a.l.ld + c]);
    float d = p - a.l.rgb[(b * 3 + c) * per + rem];
    sum_d2 += d * d;
    dl[c] = k_rgb * d * m * m * p * (1.f - p);
  }
  for (int j = 3; j < a.dld; ++j) dl[j] = 0.f;
  if (dlb) {
    __nv_bfloat162 p01 = __floats2bfloat162_rn(dl[0], dl[1]), p2 = __floats2bfloat162_rn(dl[2], 0.f);
    dlb[0] = make_uint4(*reinterpret_cast<uint32_t*>(&p01), *reinterpret_cast<uint32_t*>(&p2), 0u, 0u);
  }
  if (dm) {
    // d all / d m_p  
    double gm = (double)a.c_rgb * (2.0 * m * sum_d2 * cf.inv_n_rgb - cf.s_over_n2)
              + (double)a.c_mask * (-2.0 * (1.0 - m) * cf.inv_n_mask);
    if (a.edge_pred) {
      double e2 = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double l = a.edge_label[(b * a.label_channels + (a.label_channels == 1 ? 0 : c)) * per + rem];
        double e = a.edge_pred[(b * 3 + c) * per + rem] - l;
        e2 += e * e;
      }
      gm += (double)a.c_edge * (2.0 * m * e2 * cf.inv_n_edge - cf.se_over_n2);
    }
    dm[0] = (float)(gm * (double)(m * (1.f - m)));
    for (int j = 1; j < a.dmld; ++j) dm[j] = 0.f;
    if (dmb) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(dm[0], 0.f);
      dmb[0] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), 0u, 0u, 0u);
    }
  }
}

// plain sigmoid of the first 3 columns (render path)
static __global__ void k_sigmoid_out(int n, const float* __restrict__ logits, int ld, float* __restrict__ out) {
  pdl_wait();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[(size_t)t * 3 + c] = sigmoidf_acc(logits[(size_t)t * ld + c]);
}

// ============================================================================================
// edge branch (inputs.py:50-69): Sobel-3 x/y in fp64 -> magnitude -> 5x5 Gaussian (sigma=0 -> [1 4 6 4 1]/16),
// OpenCV default border BORDER_REFLECT_101.  Image layouts: planar [n,c,rows,w] or interleaved [n,rows*w,c].
// ============================================================================================
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

static __global__ void k_sobel_mag(const float* __restrict__ img, int n, int ch, int rows, int w, int interleaved,
                            double* __restrict__ mag /* planar [n,ch,rows,w] */) {
  pdl_wait();
  // grid: x covers one [rows, w] plane, y = plane (image * ch + channel): 32-bit index arithmetic only
  const int pi = blockIdx.x * blockDim.x + threadIdx.x;
  if (pi >= rows * w) return;
  const int y = pi / w, x = pi - y * w;
  const int c = (int)(blockIdx.y % (unsigned)ch), b = (int)(blockIdx.y / (unsigned)ch);
  const long long i = (long long)blockIdx.y * rows * w + pi;
  (void)n;
  auto px = [&](int yy, int xx) -> double {
    yy = reflect101(yy, rows); xx = reflect101(xx, w);
    return interleaved ? (double)img[(((long long)b * rows + yy) * w + xx) * ch + c]
                       : (double)img[(((long long)b * ch + c) * rows + yy) * w + xx];
  };
  double gx = 0, gy = 0;
  const double s[3] = {1, 2, 1}, d[3] = {-1, 0, 1};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      double v = px(y + dy - 1, x + dx - 1);
      gx += s[dy] * d[dx] * v;
      gy += d[dy] * s[dx] * v;
    }
  mag[i] = sqrt(gx * gx + gy * gy);
}

// Sobel magnitude -> 5x5 Gaussian -> edge statistics of one step in ONE launch (replaces k_sobel_mag + k_gauss5 + k_edge_stats
// on the training path; same fp64 operation order per pixel, so edge_pred is bit-identical).  Block = 32x8 output pixels of one
// channel of one image: the 38x14 input window (REFLECT_101 resolved while loading) and the 36x12 magnitude window live in SMEM.
// A magnitude outside the image is the magnitude AT the reflected position, as OpenCV's two-pass evaluation gives.
constexpr int kEfW = 32, kEfH = 8;
static __global__ void __launch_bounds__(256) k_edge_fused(const float* __restrict__ pred /* [batch, rows*w, 3] */, int rows, int w, EdgeArgs a,
                                                           double* __restrict__ edge_out /* [batch,3,rows,w] */,
                                                           double* __restrict__ sums) {
  pdl_wait();
  __shared__ double s_in[kEfH + 6][kEfW + 6];
  __shared__ double s_mag[kEfH + 4][kEfW + 4];
  __shared__ double red[32];
  const int b = blockIdx.z / 3, c = blockIdx.z - 3 * b;      // one (image, channel) per block: three times the parallelism
  const int x0 = blockIdx.x * kEfW, y0 = blockIdx.y * kEfH;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int x = x0 + tx, y = y0 + ty;
  const bool inside = x < w && y < rows;
  const long long per = (long long)rows * w;
  const long long pix = (long long)y * w + x;
  double m = 1.0;
  if (inside) {
    if (a.mask_mode == MARF_MASK_DISK) m = (double)a.masks_eroded[b * per + pix];
    else if (a.mask_mode == MARF_MASK_IMPLICIT) m = (double)a.mask_pred[b * per + pix];
  }
  double s_acc = 0.0;
  const double sk[3] = {1, 2, 1}, dk[3] = {-1, 0, 1};
  const double g5[5] = {1.0 / 16, 4.0 / 16, 6.0 / 16, 4.0 / 16, 1.0 / 16};
  {
    for (int i = threadIdx.x; i < (kEfH + 6) * (kEfW + 6); i += 256) {
      const int iy = i / (kEfW + 6), ix = i - iy * (kEfW + 6);
      const int yy = reflect101(y0 - 3 + iy, rows), xx = reflect101(x0 - 3 + ix, w);
      s_in[iy][ix] = (double)pred[((long long)b * per + (long long)yy * w + xx) * 3 + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kEfH + 4) * (kEfW + 4); i += 256) {
      const int my = i / (kEfW + 4), mx = i - my * (kEfW + 4);
      // magnitude at the reflected position q of window coordinate p; q's neighbours are within the loaded window
      const int qy = reflect101(y0 - 2 + my, rows) - (y0 - 3), qx = reflect101(x0 - 2 + mx, w) - (x0 - 3);
      double gx = 0, gy = 0;
      if (qy >= 1 && qy <= kEfH + 4 && qx >= 1 && qx <= kEfW + 4) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const double v = s_in[qy + dy - 1][qx + dx - 1];
            gx += sk[dy] * dk[dx] * v;
            gy += dk[dy] * sk[dx] * v;
          }
      }
      s_mag[my][mx] = sqrt(gx * gx + gy * gy);
    }
    __syncthreads();
    if (inside) {
      double acc = 0;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) {
        double rowacc = 0;
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) rowacc += g5[dx] * s_mag[ty + dy][tx + dx];
        acc += g5[dy] * rowacc;
      }
      edge_out[((long long)b * 3 + c) * per + pix] = acc;
      const double l = a.edge_label[((long long)b * a.label_channels + (a.label_channels == 1 ? 0 : c)) * per + pix];
      const double d = (acc - l) * m;
      s_acc += d * d;
    }
  }
  block_sum_atomic(s_acc, &sums[MARF_S_EDGE], red);
  block_sum_atomic(inside ? m : 0.0, &sums[MARF_N_EDGE], red);        // (the three channel blocks of a pixel add m each: 3 m)
}

static __global__ void k_gauss5(const double* __restrict__ mag, int planes, int rows, int w, double* __restrict__ out) {
  pdl_wait();
  const int pi = blockIdx.x * blockDim.x + threadIdx.x;     // grid: x covers one plane, y = plane
  if (pi >= rows * w) return;
  const int y = pi / w, x = pi - y * w;
  const long long p = blockIdx.y;
  const long long i = p * rows * w + pi;
  (void)planes;
  const double g5[5] = {1.0 / 16, 4.0 / 16, 6.0 / 16, 4.0 / 16, 1.0 / 16};
  const double* pl = mag + p * rows * w;
  // separable, rows first then columns (same result as OpenCV's row/column filter order up to fp64 rounding)
  double acc = 0;
#pragma unroll
  for (int dy = 0; dy < 5; ++dy) {
    int yy = reflect101(y + dy - 2, rows);
    double rowacc = 0;
#pragma unroll
    for (int dx = 0; dx < 5; ++dx) rowacc += g5[dx] * pl[(long long)yy * w + reflect101(x + dx - 2, w)];
    acc += g5[dy] * rowacc;
  }
  out[i] = acc;
}

}  // namespace marf
