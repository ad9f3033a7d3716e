// fp32-parity GEMMs on the tensor cores: 3xTF32 (sm_100a, tcgen05.mma kind::tf32 on CTA pairs).
//
// Every fp32 operand x is split into  big = tf32(x)  and  small = tf32(x - big)  (round to nearest, 11 + 11 significant bits) and a
// product is evaluated as  small_a*big_b + big_a*small_b + big_a*big_b  with fp32 accumulation in TMEM.  Measured against
// float64 (profiles/tools/tf32_probe.py, K = 256): relative L2 error 1.8e-6 — fp32 SGEMM 2.9e-7, one-pass TF32 2.9e-4; the
// remainder is the tensor core's truncating fp32 accumulate, which grows with K.  Rate: 1/3 of TF32 = 1/6 of bf16, instead of
// the CUDA-core rate.  These kernels replace k_sgemm (fp32_kernels.cuh) for the wide layers of precision=fp32; k_sgemm stays
// for the narrow ones (N or K < 32) and as the reference implementation in tests (MARF_FP32_TC=0).
//
//   k_tf32x3<MODE_NT, EPI> : C[M,N] = epi(A[M,K] * B[N,K]^T)        forward layers (B = W) and dX layers (B = W^T)
//   k_tf32x3<MODE_TN, *>   : C[N,K'] += P[M,N]^T * Q[M,K'] (+ db)    dW layers (contraction over the pixel rows), split over
//                                                                    row ranges, red.add at the end
//
// CTA pairs (clusters of 2, cta_group::2): one MMA is M = 256 (128 accumulator rows in each CTA's TMEM) x N <= 256 with the B
// operand split across the pair (N/2 rows in each CTA's SMEM).  The shape follows from the SMEM pipe, which is what bounds a
// 3-term product: a single-CTA M = 128 x N = 256 MMA re-reads 12 KB of operands per 128 cycles (96 of the 128 B/clk) while
// the stage's 96 KB of planes have to be written — measured 3,000 cycles per stage, 2x the MMA time; in the pair a CTA reads
// 8 KB per MMA, writes 64 KB per stage and has room for a ring of three stages.
//
// How the planes get into SMEM (history and measurements in profiles/r02_tf32_gemm.txt):
//   weights (MODE_NT B):   split once per step (k_tf32_split_w), then 2 TMA boxes per stage and CTA (SWIZZLE_128B = the
//                          K-major operand layout): no thread touches them
//   activations (MODE_NT A): 8 loader warps, 16-byte chunks global -> registers (two register sets = two stages in flight)
//                          -> two integer roundings (cvt.rna.tf32 issues at a quarter rate) -> st.shared.v4 of both planes
//   MODE_TN (P and Q, both need a transposition: the contraction runs over rows): raw [32 x 128] tiles by TMA into a ring of
//                          two raw stages, conflict-free ld.shared.v4, 4-byte scatter into the planes with a lane mapping
//                          (16 rows x 2 chunks per warp) that hits 32 distinct banks; column sums of P (the bias gradient)
//                          accumulate in registers on the way.  (Per-thread global loads of that shape are 16 x 32-byte
//                          sectors per instruction and cost 4,100 cycles per stage; with TMA the kernel runs at the MMA rate.)
//   epilogue (MODE_NT):    thread = accumulator row; bias / ReLU / sign bits / bit mask on the TMEM layout, [32 x 32] blocks
//                          through a swizzled staging buffer and one TMA store each.  The forward layers leave one sign bit per
//                          output (T_BIAS_RELU bits_out); the dX layers mask with those bits (T_RELU_BITS) instead of re-reading
//                          the fp32 layer input (T_RELU_MASK, kept for inputs that did not come from a tensor-core layer).
//
// Warp roles, MODE_NT (416 threads): warp 0 = MMA issuer (leader CTA; + TMEM alloc), warps 1..8 = loaders, warps 9..12 =
// epilogue (warp w owns TMEM lanes 32*(w%4)..+31); MODE_TN (384 threads): warps 0..7 = loaders, warps 8..11 = epilogue, lane
// 0 of warp 8 issues the MMAs first and lane 0 of warp 9 the TMA loads.  SMEM stages of 32 K-elements (A and B-half: 2
// planes x 16 KB each), two TMEM accumulators of 256 columns.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace marf {
namespace t32 {

using namespace marf::tc;

constexpr int kThreads = 416;             // MODE_NT
constexpr int kThreadsTN = 384;           // MODE_TN: the MMA issuer is lane 0 of the first epilogue warp (the epilogue starts when the
                                          // MMAs are done), 12 warps leave 168 registers for the two register sets of the loaders
constexpr int kTileM = 128;               // accumulator rows per CTA (the pair's MMA is M = 256)
constexpr int kStageK = 32;               // tf32 elements per 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kPlane = kTileM * 128;      // 16 KB: one [128 rows x 32] plane
constexpr int kStageBytes = 4 * kPlane;   // A big, A small, B-half big, B-half small
constexpr int kOutStage = 8 * 4096;       // epilogue staging: two [32 rows x 32 columns] blocks per warp
// MODE_TN: two plane stages + a ring of two raw stages: the [32 rows x 128 columns] tiles of P and of this CTA's half of Q as
// they lie in memory, brought in by TMA as 4 + 4 boxes of [32 rows x 32 columns] (SWIZZLE_128B, so that the 8 consecutive rows
// of one 16-byte column chunk a quarter-warp reads fall into 8 different bank groups)
constexpr int kRawBox = kStageK * 128;               // 4 KB
constexpr int kRawStage = 8 * kRawBox;               // 32 KB
constexpr int kRawOff = 2 * kStageBytes;
constexpr int kSmemBytes = kStages * kStageBytes + kOutStage + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment*/;

enum { MODE_NT = 0, MODE_TN = 1 };
enum { T_BIAS = 0, T_BIAS_RELU = 1, T_PLAIN = 2, T_RELU_MASK = 3, T_RELU_BITS = 4 };

struct Params {
  // MODE_NT: C[M, 0..n_valid) = epi(A[M,K] * B[n_valid,K]^T);  A row-major [M, lda]; B = the weights, split once per step into
  //          big / small planes (k_tf32_split_w) and brought in by TMA (tmB);
  //          pair i works on the 256-row super-tiles i, i + pairs, ...; CTA r of the pair owns rows r*128.. of a super-tile
  //          and stages rows r*npad/2.. of B
  // MODE_TN: C[p, q] += sum_m A[m, p] * B[m, q];  A = P [M, lda], B = Q [M, ldb];  db[p] += sum_m P[m, p];
  //          pair y takes 256 columns of P (CTA r: 128 of them), pair z 256 columns of Q (CTA r stages half of them),
  //          pair x a range of rows
  const float* A;
  int lda;
  const float* B;
  int ldb;
  float* C;
  int ldc;
  const float* aux;      // bias [N] (T_BIAS*), or the layer input [M, ldaux] (T_RELU_MASK)
  int ldaux;
  float* db;             // MODE_TN: column sums of P (nullptr: not wanted)
  uint32_t* bits_out;    // T_BIAS_RELU: optional [M, bits_ld] words, bit c of word g of a row = (output column 32 g + c > 0)
  const uint32_t* bits_in;   // T_RELU_BITS: the same layout, written by the forward layer that produced this layer's input
  int bits_ld;           // words per row
  int M;                 // rows (multiple of 128)
  int K;                 // MODE_NT: contraction length
  int n_valid;           // MODE_NT: output columns (<= 256);  MODE_TN: columns of Q
  int p_valid;           // MODE_TN: columns of P
  int splits;            // MODE_TN: pairs along the rows
  CUtensorMap tmB;       // MODE_NT: the split weights [2 x rows, K padded to 32] (big planes, then small planes from row b_small),
                         //          box [bh rows x 32 columns], SWIZZLE_128B
  int b_row0, b_small;   // MODE_NT: first row of this launch's column chunk; row offset of the small planes
  CUtensorMap tmC;       // MODE_NT: fp32 [M rows, n_valid columns] map of C, box [32 x 32], SWIZZLE_128B (TMA stores clip)
  CUtensorMap tmP, tmQ;  // MODE_TN: fp32 [M rows, valid columns] maps of P and Q, box [32 x 32], SWIZZLE_128B, zero fill
  long long* trace;      // diagnostics (MARF_T32_TRACE): clock64() stamps of CTA 0, [stage or tile][8]; nullptr in production
};

// canonical K-major SWIZZLE_128B position of element (row, k) of a [rows x 32] fp32 plane (bytes)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 2) ^ (row & 7u)) & 7u) << 4) + (k & 3u) * 4u;
}

// round to the nearest tf32 (ties away from zero), as cvt.rna.tf32.f32 does for finite values — with two integer
// instructions: F2F conversions issue at a quarter of the rate and 96 of them per thread and stage bound the loaders
__host__ __device__ __forceinline__ float tf32_rna(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
#endif
}

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same on a CTA pair: D[256 rows: 128 in each CTA's TMEM] (+)= A[128 rows from each CTA] * B[N/2 rows from each CTA]
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// streaming accesses of the epilogue: every line is touched once, keep them out of L1
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// explicit shared-space accesses (32-bit shared addresses): through the rounded-up dynamic-SMEM pointer the compiler only
// sees generic pointers and emits ST.E / LD.E, which cost the loaders 2-3x the issue time of STS / LDS
__device__ __forceinline__ void sts128(uint32_t a, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float4 ldg128(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void split_store4(uint32_t big, uint32_t small, const float4& v) {
  float4 b, s;
  b.x = tf32_rna(v.x); b.y = tf32_rna(v.y); b.z = tf32_rna(v.z); b.w = tf32_rna(v.w);
  s.x = tf32_rna(v.x - b.x); s.y = tf32_rna(v.y - b.y); s.z = tf32_rna(v.z - b.z); s.w = tf32_rna(v.w - b.w);
  sts128(big, b);
  sts128(small, s);
}

// Weights -> big / small tf32 planes, once per step:  out[n][k] = big(B[n][k]),  out[nr + n][k] = small(B[n][k]),  [2 nr, kpad]
// row-major, zero outside (n < N, k < K).   TRANS = 0: B[n][k] = W[n * ldw + k] (forward: n = out, k = in);
// TRANS = 1: B[n][k] = W[k * ldw + n] (dX: n = in, k = out)
template <int TRANS>
static __global__ void k_tf32_split_w(const float* __restrict__ W, int ldw, int N, int K, float* __restrict__ out, int nr, int kpad) {
  pdl_wait();
  const int total = nr * kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int n, k;
    if (TRANS) { n = i % nr; k = i / nr; }      // consecutive threads read consecutive n (contiguous in W)
    else       { k = i % kpad; n = i / kpad; }
    float x = 0.f;
    if (n < N && k < K) x = TRANS ? W[(size_t)k * ldw + n] : W[(size_t)n * ldw + k];
    const float big = tf32_rna(x);
    out[(size_t)n * kpad + k] = big;
    out[(size_t)(nr + n) * kpad + k] = tf32_rna(x - big);
  }
}

template <int MODE, int EPI>
__global__ void __launch_bounds__(MODE == MODE_TN ? kThreadsTN : kThreads, 1) k_tf32x3(const __grid_constant__ Params p) {
  constexpr int kLoaderWarp0 = MODE == MODE_TN ? 0 : 1;
  constexpr int kEpiWarp0 = kLoaderWarp0 + 8;
  constexpr int kMmaWarp = MODE == MODE_TN ? kEpiWarp0 : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sOut = smem + kStages * kStageBytes;
  float* sBias = reinterpret_cast<float*>(sOut + kOutStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + kOutStage + 1024);
  uint64_t* full = bars;                    // [kStages]  leader CTA: 16 loader-warp arrivals (8 local + 8 from the peer)
  uint64_t* empty = bars + kStages;         // [kStages]  both CTAs: MMAs of the stage retired (multicast commit)
  uint64_t* acc_full = bars + 2 * kStages;  // [2]        both CTAs (multicast commit)
  uint64_t* acc_empty = acc_full + 2;       // [2]        leader CTA: 8 epilogue-warp arrivals
  uint64_t* raw_full = acc_empty + 2;       // [2]        MODE_TN: bytes of a raw stage landed
  uint64_t* raw_empty = raw_full + 2;       // [2]        MODE_TN: 8 loader warps have read the raw stage
  uint64_t* full_b = raw_empty + 2;         // [kStages]  MODE_NT, leader CTA: the weight planes of both CTAs landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_b + kStages);
  constexpr int kSt = MODE == MODE_TN ? 2 : kStages;      // plane stages
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0 = leader (issues the MMAs)
  const int pair = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;

  // ---- work of this pair
  //   MODE_NT: 256-row super-tiles pair, pair + n_pairs, ... each over ceil(K / 32) stages
  //   MODE_TN: one accumulator; a contiguous range of the 32-row stages
  int n_tiles_my, st_begin = 0, st_count;
  if (MODE == MODE_NT) {
    const int n_super = (p.M / kTileM + 1) / 2;
    n_tiles_my = pair < n_super ? (n_super - pair + n_pairs - 1) / n_pairs : 0;
    st_count = (p.K + kStageK - 1) / kStageK;
  } else {
    const int st_total = p.M / kStageK;
    const int per = (st_total + p.splits - 1) / p.splits;
    st_begin = pair * per;
    st_count = max(0, min(per, st_total - st_begin));
    n_tiles_my = st_count > 0 ? 1 : 0;
  }
  const int nv = MODE == MODE_TN ? min(256, p.n_valid - (int)blockIdx.z * 256) : p.n_valid;   // valid accumulator columns
  const int npad = (nv + 15) / 16 * 16;                                                       // MMA N
  const int bh = npad / 2;                                   // B-operand rows staged by each CTA
  const int p0 = MODE == MODE_TN ? (int)blockIdx.y * 256 + (int)rank * 128 : 0;   // first column of P (= row of C) of this CTA
  const int q0 = MODE == MODE_TN ? (int)blockIdx.z * 256 : 0;                    // first column of Q (= column of C) of the pair
  const int pv = MODE == MODE_TN ? max(0, min(128, p.p_valid - p0)) : 128;       // valid accumulator rows of this CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 16); mbar_init(&empty[s], 1); mbar_init(&full_b[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 8); mbar_init(&raw_full[a], 1); mbar_init(&raw_empty[a], 8); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  pdl_wait();                                // everything below reads the previous kernels' output
  tc_fence_before();
  cluster_sync_all();                        // the peer's barriers must be initialised before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_it = n_tiles_my * st_count;    // stages of this pair

  if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = idesc_tf32(2 * kTileM, npad);
      uint32_t it = 0;
      for (int t = 0; t < n_tiles_my; ++t) {
        const uint32_t a = t & 1, aph = (t >> 1) & 1;
        mbar_wait_cluster(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * 256;
        for (int ks = 0; ks < st_count; ++ks, ++it) {
          const uint32_t s = it % kSt, ph = (it / kSt) & 1;
          mbar_wait_cluster(&full[s], ph);
          if (MODE == MODE_NT) mbar_wait_cluster(&full_b[s], ph);
          tc_fence_after();
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && it < 512) p.trace[it * 8 + 4] = clock64();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint64_t da_big = smem_desc_sw128(sa, 16, 1024);
          const uint64_t da_small = smem_desc_sw128(sa + kPlane, 16, 1024);
          const uint64_t db_big = smem_desc_sw128(sa + 2 * kPlane, 16, 1024);
          const uint64_t db_small = smem_desc_sw128(sa + 3 * kPlane, 16, 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) {       // 4 K-steps of 8 tf32 (32 bytes) per stage; small terms first
            umma_tf32_2sm(d_tmem, da_small + 2 * j, db_big + 2 * j, idesc, (ks | j) != 0);
            umma_tf32_2sm(d_tmem, da_big + 2 * j, db_small + 2 * j, idesc, 1);
            umma_tf32_2sm(d_tmem, da_big + 2 * j, db_big + 2 * j, idesc, 1);
          }
          umma_commit_2sm(&empty[s]);
          if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && it < 512) p.trace[it * 8 + 5] = clock64();
        }
        umma_commit_2sm(&acc_full[a]);
      }
    }
    __syncwarp();
  }
  if (warp >= kLoaderWarp0 && warp < kEpiWarp0) {
    // ------------------------------------------------------------------ loaders: split (and transpose) into the stages
    const int lt = threadIdx.x - 32 * kLoaderWarp0;          // 0..255
    const int lw = lt >> 5;                   // loader warp 0..7
    // two register sets = the operands of two stages in flight: the loads of stage i + 2 are issued as soon as the registers
    // of stage i have been written to SMEM (they do not wait for an SMEM slot)
    float4 va[4], vb[4];
    const uint32_t full_remote = mapa_u32(smem_u32(&full[0]), 0);      // the leader's full[] barriers
    auto publish = [&](uint32_t s) {
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(full_remote + 8u * s);
    };
    if (MODE == MODE_NT) {
      // A: chunk (row0 + 32 j, kc), row0 = lt >> 3 (0..31), kc = lt & 7, j = 0..3 (this CTA's 128 rows); 8 lanes read one
      // 128-byte row segment, SW128 offset of (row0 + 32 j, kc) = base + 4096 j.  B (the pre-split weights): two TMA boxes per
      // stage and CTA, issued by the first loader thread as soon as the slot is free.
      const int row0 = lt >> 3, kc = lt & 7;
      const uint32_t base = (uint32_t)(row0 >> 3) * 1024u + (uint32_t)(row0 & 7) * 128u + (uint32_t)((kc ^ (row0 & 7)) << 4);
      auto issue = [&](float4 (&v)[4], int i) {
        if (i >= n_it) return;
        const int t = i / st_count, ks = i - t * st_count;
        const int row = ((pair + t * n_pairs) * 2 + (int)rank) * kTileM + row0;      // first of this thread's 4 rows of A
        const int k = ks * kStageK + kc * 4;
        const bool ain = k < p.K && row < p.M;             // (the last super-tile may have no second half)
        const float* gA = p.A + (size_t)row * p.lda + k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ain) v[j] = ldg128(gA + (size_t)(32 * j) * p.lda);
        }
      };
      auto stage = [&](float4 (&v)[4], int i) {
        if (i >= n_it) return;
        const uint32_t s = (uint32_t)i % kStages, ph = ((uint32_t)i / kStages) & 1;
        mbar_wait_cluster(&empty[s], ph ^ 1);
        if (lt == 0) {
          const int ks = i % st_count;
          uint8_t* sb = smem + s * kStageBytes + 2 * kPlane;
          if (rank == 0) mbar_expect_tx(&full_b[s], 4u * (uint32_t)bh * 128u);
          tma_load_2d_2sm(sb, &p.tmB, ks * kStageK, p.b_row0 + (int)rank * bh, &full_b[s], kEvictLast);
          tma_load_2d_2sm(sb + kPlane, &p.tmB, ks * kStageK, p.b_small + p.b_row0 + (int)rank * bh, &full_b[s], kEvictLast);
        }
        const bool tr = p.trace && blockIdx.x == 0 && lt == 0 && i < 512;
        if (tr) p.trace[i * 8 + 0] = clock64();
        const uint32_t sa = smem_u32(smem) + s * kStageBytes + base;
#pragma unroll
        for (int j = 0; j < 4; ++j) split_store4(sa + j * 4096, sa + kPlane + j * 4096, v[j]);
        if (tr) p.trace[i * 8 + 1] = clock64();
        publish(s);
        if (tr) p.trace[i * 8 + 2] = clock64();
        issue(v, i + 2);
      };
      if (lt == 0) prefetch_tmap(&p.tmB);
      issue(va, 0);
      issue(vb, 1);
      for (int i = 0; i < n_it; i += 2) {
        stage(va, i);
        stage(vb, i + 1);
      }
    } else {
      // thread -> row m of the stage and column chunk c4_0 (of 4 floats); unit j adds 8 chunks (32 columns): P for j < 4
      // (this CTA's 128 columns), Q for j = 4..7 (this CTA's half of the pair's columns).  Component c of a chunk goes to
      // operand row 4 c4 + c, element m: SW128 offset = base[c] + 4096 j.  Within a warp m covers 16 values and c4 two
      // (lane >> 4): rows 4 c4 + c of the two halves differ by 4 in their swizzle phase, so the 32 lanes hit 32 distinct
      // banks; the loads are 16 x 32-byte sectors.
      const int m = (lw & 1) * 16 + (lane & 15);
      const int c4_0 = (lw >> 1) * 2 + (lane >> 4);      // 0..7
      uint32_t base[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) base[c] = sw128_off((uint32_t)(4 * c4_0 + c), (uint32_t)m);
      const int qc0 = q0 + (int)rank * bh;               // first column of Q staged by this CTA
      int nP = 0, nQ = 0, nQs = 0;             // units inside the matrices / read by the MMA (operand rows < bh)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        nP += 4 * c4_0 + 32 * j < pv ? 1 : 0;
        nQ += (4 * c4_0 + 32 * j < bh && qc0 + 4 * c4_0 + 32 * j < p.n_valid) ? 1 : 0;
        nQs += 4 * c4_0 + 32 * j < bh ? 1 : 0;
      }
      float colsum[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) colsum[i] = 0.f;
      const uint32_t raw0 = smem_u32(smem) + kRawOff + (uint32_t)m * 128u + (uint32_t)((c4_0 ^ (m & 7)) << 4);
      for (int i = 0; i < n_it; ++i) {
        // raw stage -> registers (conflict-free 16-byte reads), release the raw slot, then wait for a plane slot
        const uint32_t rs = (uint32_t)i & 1u, rph = ((uint32_t)i >> 1) & 1u;
        mbar_wait(&raw_full[rs], rph);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nP) v[j] = lds128(raw0 + rs * kRawStage + j * kRawBox);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[4 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nQ) v[4 + j] = lds128(raw0 + rs * kRawStage + (4 + j) * kRawBox);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&raw_empty[rs]);
        const uint32_t s = (uint32_t)i % kSt, ph = ((uint32_t)i / kSt) & 1;
        mbar_wait_cluster(&empty[s], ph ^ 1);
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lt == 0 && i < 512;
        if (tr) p.trace[i * 8 + 0] = clock64();
        const uint32_t sa = smem_u32(smem) + s * kStageBytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j >= 4 && j - 4 >= nQs) continue;
          const uint32_t pl = sa + (j < 4 ? j * 4096 : 2 * kPlane + (j - 4) * 4096);
          const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float b = tf32_rna(x[c]);
            sts32(pl + base[c], b);
            sts32(pl + kPlane + base[c], tf32_rna(x[c] - b));
            if (j < 4) colsum[j * 4 + c] += x[c];
          }
        }
        if (tr) p.trace[i * 8 + 1] = clock64();
        publish(s);
        if (tr) p.trace[i * 8 + 2] = clock64();
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lt == 224 && i < 512) p.trace[i * 8 + 3] = clock64();   // last loader warp
        if (p.trace && blockIdx.x == 1 && blockIdx.y == 0 && lt == 0 && i < 512) p.trace[i * 8 + 6] = clock64();     // peer CTA
      }
      if (p.db && blockIdx.z == 0 && st_count > 0) {
        // column sums of P: reduce over the 16 rows m of this warp (lane bits 0..3), one atomic per column and warp
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float sum = colsum[i];
#pragma unroll
          for (int o = 1; o < 16; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const int col = 4 * c4_0 + 32 * (i >> 2) + (i & 3);
          if ((lane & 15) == 0 && col < pv) atomicAdd(p.db + p0 + col, sum);
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    if (MODE == MODE_TN && warp == kEpiWarp0 + 1 && lane == 0) {
      // ---------------------------------------------------------------- raw-stage producer (MODE_TN): 8 TMA boxes per stage.
      // (row-wise cp.async.bulk copies — 64 per stage — cost the copy engine ~65 cycles each: 4,200 cycles per stage)
      const int qc0 = q0 + (int)rank * bh;
      prefetch_tmap(&p.tmP);
      prefetch_tmap(&p.tmQ);
      for (int i = 0; i < n_it; ++i) {
        const uint32_t rs = (uint32_t)i & 1u, rph = ((uint32_t)i >> 1) & 1u;
        mbar_wait(&raw_empty[rs], rph ^ 1);
        mbar_expect_tx(&raw_full[rs], kRawStage);
        uint8_t* dst = smem + kRawOff + rs * kRawStage;
        const int r0 = (st_begin + i) * kStageK;
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(dst + j * kRawBox, &p.tmP, p0 + 32 * j, r0, &raw_full[rs]);
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(dst + (4 + j) * kRawBox, &p.tmQ, qc0 + 32 * j, r0, &raw_full[rs]);
      }
    }
    __syncwarp();
    // ------------------------------------------------------------------ epilogue (4 warps per CTA)
    const int q = warp & 3;
    const uint32_t stg = smem_u32(sOut) + q * 4096;   // this warp's [32 rows x 32 columns] staging block, 16-byte chunks swizzled by row
    if ((EPI == T_BIAS || EPI == T_BIAS_RELU) && MODE == MODE_NT) {
      for (int i = threadIdx.x - 32 * kEpiWarp0; i < 256; i += 128) sts32(smem_u32(sBias) + 4u * i, i < nv ? p.aux[i] : 0.f);
      named_bar_sync(1, 128);
    }
    const uint32_t acc_empty_remote = mapa_u32(smem_u32(&acc_empty[0]), 0);
    const int n_groups = (npad + 31) / 32;
    auto tile_row0 = [&](int t) { return ((pair + t * n_pairs) * 2 + (int)rank) * kTileM; };
    constexpr bool kDirect = MODE == MODE_NT && EPI != T_RELU_MASK;
    if (kDirect) {
      // Thread = accumulator row: bias / ReLU / sign bits / bit mask are applied on the TMEM row layout, the [32 x 32] block of the
      // warp goes to a swizzled staging buffer and from there to global memory by one TMA store (two buffers per warp: the store
      // of a column group drains while the next one is prepared).  No read-back, no per-thread global stores.
      uint8_t* buf_ptr = sOut + q * 8192;
      const uint32_t buf0 = smem_u32(buf_ptr);
      uint32_t n_buf = 0;
      auto load_word = [&](int t, int g) -> uint32_t {
        if (EPI != T_RELU_BITS || t >= n_tiles_my) return 0u;
        const int row = tile_row0(t) + q * 32 + lane;
        return row < p.M ? p.bits_in[(size_t)row * p.bits_ld + g] : 0u;
      };
      uint32_t wnext = load_word(0, 0);
      if (lane == 0) prefetch_tmap(&p.tmC);
      for (int t = 0; t < n_tiles_my; ++t) {
        const uint32_t a = t & 1, aph = (t >> 1) & 1;
        mbar_wait_cluster(&acc_full[a], aph);
        tc_fence_after();
        const bool tre = p.trace && blockIdx.x == 0 && threadIdx.x == 32 * kEpiWarp0 && t < 64;
        if (tre) p.trace[t * 8 + 6] = clock64();
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256;
        const int row0 = tile_row0(t) + q * 32;
        const bool rows_in = row0 < p.M;
        uint32_t v[32];
        tmem_ld32(tbase, v);
        for (int g = 0; g < n_groups; ++g, ++n_buf) {
          const uint32_t boff = (n_buf & 1u) * 4096u;
          if (lane == 0 && n_buf >= 2) bulk_wait_read<1>();        // the store that read this buffer two groups ago is done
          __syncwarp();
          const uint32_t w = wnext;
          wnext = g + 1 < n_groups ? load_word(t, g + 1) : load_word(t + 1, 0);
          tmem_ld_wait();
          uint32_t ob = 0;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 y = make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]), __uint_as_float(v[c4 * 4 + 2]),
                                   __uint_as_float(v[c4 * 4 + 3]));
            if (EPI == T_BIAS || EPI == T_BIAS_RELU) {
              const float4 b = lds128(smem_u32(sBias) + 4u * (g * 32 + c4 * 4));
              y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
              if (EPI == T_BIAS_RELU) {
                y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
                ob |= ((y.x > 0.f ? 1u : 0u) | (y.y > 0.f ? 2u : 0u) | (y.z > 0.f ? 4u : 0u) | (y.w > 0.f ? 8u : 0u)) << (4 * c4);
              }
            } else if (EPI == T_RELU_BITS) {
              const uint32_t b4 = w >> (4 * c4);
              y.x = (b4 & 1u) ? y.x : 0.f; y.y = (b4 & 2u) ? y.y : 0.f; y.z = (b4 & 4u) ? y.z : 0.f; y.w = (b4 & 8u) ? y.w : 0.f;
            }
            sts128(buf0 + boff + lane * 128 + ((c4 ^ (lane & 7)) << 4), y);
          }
          if (g + 1 < n_groups) tmem_ld32(tbase + (g + 1) * 32, v);     // next group's accumulator columns travel meanwhile
          if (EPI == T_BIAS_RELU && p.bits_out && rows_in && g < p.bits_ld) p.bits_out[(size_t)(row0 + lane) * p.bits_ld + g] = ob;
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && rows_in) {
            tma_store_2d(&p.tmC, g * 32, row0, buf_ptr + boff);
            bulk_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_remote + 8u * a);
        if (tre) p.trace[t * 8 + 7] = clock64();
      }
      if (lane == 0) bulk_wait<0>();
    } else {
    const int rr = lane >> 3, ch = lane & 7;   // read-back: rows rr + 4 i, chunk ch -> 8 lanes cover one 128-byte line
    // T_RELU_MASK: the mask (the layer input) does not depend on the accumulator: the 8 lines a thread needs for a column
    // group are requested one group ahead (the first ones before the accumulator is even complete)
    float4 mk[EPI == T_RELU_MASK ? 8 : 1];
    uint32_t mw[8];                            // T_RELU_BITS: the mask words of rows rr + 4 i
    auto load_mask = [&](int t, int g) {
      if (EPI == T_RELU_BITS && MODE == MODE_NT && t < n_tiles_my) {
        const int grow0 = tile_row0(t) + q * 32 + rr;
#pragma unroll
        for (int i = 0; i < 8; ++i) mw[i] = grow0 < p.M ? p.bits_in[(size_t)(grow0 + 4 * i) * p.bits_ld + g] : 0u;
      }
      if (EPI != T_RELU_MASK || MODE != MODE_NT || t >= n_tiles_my) return;
      const int col = g * 32 + ch * 4;
      const int grow0 = tile_row0(t) + q * 32 + rr;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        mk[i] = (col < nv && grow0 < p.M) ? ldg_stream(p.aux + (size_t)(grow0 + 4 * i) * p.ldaux + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load_mask(0, 0);
    for (int t = 0; t < n_tiles_my; ++t) {
      const uint32_t a = t & 1, aph = (t >> 1) & 1;
      mbar_wait_cluster(&acc_full[a], aph);
      tc_fence_after();
      const bool tre = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 32 * kEpiWarp0 && t < 64;
      if (tre) p.trace[t * 8 + 6] = clock64();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256;
      const bool rows_in = MODE == MODE_TN || tile_row0(t) < p.M;
      uint32_t v[32];
      tmem_ld32(tbase, v);
      for (int g = 0; g < n_groups; ++g) {
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          sts128u(stg + lane * 128 + ((c4 ^ (lane & 7)) << 4), v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        if (g + 1 < n_groups) tmem_ld32(tbase + (g + 1) * 32, v);       // next group's accumulator columns travel meanwhile
        __syncwarp();
        uint32_t curbits = 0;                    // bit 4 i + c: element c of row rr + 4 i passes
        if (EPI == T_RELU_MASK) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            curbits |= ((mk[i].x > 0.f ? 1u : 0u) | (mk[i].y > 0.f ? 2u : 0u) | (mk[i].z > 0.f ? 4u : 0u) | (mk[i].w > 0.f ? 8u : 0u)) << (4 * i);
        } else if (EPI == T_RELU_BITS) {
#pragma unroll
          for (int i = 0; i < 8; ++i) curbits |= ((mw[i] >> (4 * ch)) & 15u) << (4 * i);
        }
        if (g + 1 < n_groups) load_mask(t, g + 1); else load_mask(t + 1, 0);
        const int col = g * 32 + ch * 4;
        uint32_t obits[8];                       // T_BIAS_RELU with bits_out: this thread's 4 sign bits of rows rr + 4 i
#pragma unroll
        for (int i = 0; i < 8; ++i) obits[i] = 0u;
        if (col < nv && rows_in) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rr + 4 * i;          // row inside this warp's 32
            float4 y = lds128(stg + row * 128 + ((ch ^ (row & 7)) << 4));
            if (MODE == MODE_NT) {
              const size_t grow = (size_t)(tile_row0(t) + q * 32 + row);
              if (EPI == T_BIAS || EPI == T_BIAS_RELU) {
                const float4 b = lds128(smem_u32(sBias) + 4u * col);
                y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
                if (EPI == T_BIAS_RELU) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                if (EPI == T_BIAS_RELU && p.bits_out) obits[i] = (y.x > 0.f ? 1u : 0u) | (y.y > 0.f ? 2u : 0u) | (y.z > 0.f ? 4u : 0u) | (y.w > 0.f ? 8u : 0u);
              } else if (EPI == T_RELU_MASK || EPI == T_RELU_BITS) {
                const uint32_t b4 = curbits >> (4 * i);
                y.x = (b4 & 1u) ? y.x : 0.f; y.y = (b4 & 2u) ? y.y : 0.f; y.z = (b4 & 4u) ? y.z : 0.f; y.w = (b4 & 8u) ? y.w : 0.f;
              }
              stg_stream(p.C + grow * p.ldc + col, y);
            } else {
              const int crow = q * 32 + row;
              if (crow < pv) red_add_v4(p.C + (size_t)(p0 + crow) * p.ldc + q0 + col, y.x, y.y, y.z, y.w);
            }
          }
        }
        if (EPI == T_BIAS_RELU && MODE == MODE_NT && p.bits_out) {
          // word of (row, g) = the 4-bit nibbles of the 8 lanes ch = 0..7 of that row
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint32_t w = obits[i] << (4 * ch);
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if (ch == 0 && rows_in && g < p.bits_ld) p.bits_out[(size_t)(tile_row0(t) + q * 32 + rr + 4 * i) * p.bits_ld + g] = w;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_remote + 8u * a);
      if (tre) p.trace[t * 8 + 7] = clock64();
    }
    }
  }
  tc_fence_before();
  cluster_sync_all();        // the peer may still signal this CTA's barriers / read its SMEM until here
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

}  // namespace t32
}  // namespace marf
