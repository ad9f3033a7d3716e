// fp32-parity GEMMs on the tensor cores: 3xTF32 (sm_100a, tcgen05.mma kind::tf32).
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
// Operands stay plain row-major fp32 in global memory: 8 loader warps bring 16-byte chunks into registers, split them and
// write both planes into the K-major SWIZZLE_128B layout the MMA reads (MODE_TN: transposing 4-byte scatter, lane mapping
// chosen so that it is free of bank conflicts).  Splitting the weights in the kernel instead of streaming a pre-split image
// halves the L2 -> SM bytes of a stage, which is what bounds the MODE_NT mainloop (every CTA re-reads the whole weight
// matrix for every 128-row tile).
//
// Warp roles, MODE_NT (416 threads): warp 0 = MMA issuer (+ TMEM alloc), warps 1..8 = loaders, warps 9..12 = epilogue (warp
// w owns TMEM lanes 32*(w%4)..+31); MODE_TN (384 threads): warps 0..7 = loaders, warps 8..11 = epilogue, lane 0 of warp 8
// issues the MMAs first.  Two SMEM stages of 32 K-elements (A: 2 planes x 16 KB, B: 2 planes x 32 KB), two TMEM
// accumulators of 256 columns, one 4 KB staging block per epilogue warp (the accumulator rows are re-read transposed so that
// global stores / mask loads are whole 128-byte lines).
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace marf {
namespace t32 {

using namespace marf::tc;

constexpr int kThreads = 416;             // MODE_NT
constexpr int kThreadsTN = 384;           // MODE_TN: the MMA issuer is lane 0 of the first epilogue warp (the epilogue starts when the
                                          // MMAs are done), 12 warps leave 168 registers for the two register sets of the loaders
constexpr int kTileM = 128;
constexpr int kStageK = 32;              // tf32 elements per 128-byte swizzle row
constexpr int kStages = 2;
constexpr int kPlaneA = kTileM * 128;    // 16 KB
constexpr int kPlaneB = 256 * 128;       // 32 KB
constexpr int kStageBytes = 2 * kPlaneA + 2 * kPlaneB;   // 96 KB
constexpr int kOutStage = 4 * 4096;      // epilogue staging, 4 KB per warp
constexpr int kSmemBytes = kStages * kStageBytes + kOutStage + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment*/;

enum { MODE_NT = 0, MODE_TN = 1 };
enum { T_BIAS = 0, T_BIAS_RELU = 1, T_PLAIN = 2, T_RELU_MASK = 3 };

struct Params {
  // MODE_NT: C[M, 0..n_valid) = epi(A[M,K] * B[n_valid,K]^T);  A row-major [M, lda], B row-major [n_valid, ldb]
  // MODE_TN: C[p0 + i, q0 + j] += sum_m A[m, p0 + i] * B[m, q0 + j];  A = P [M, lda], B = Q [M, ldb];
  //          CTA y takes 128 columns of P, CTA z 256 columns of Q, CTA x a range of rows;  db[p0 + i] += sum_m P[m, p0 + i]
  const float* A;
  int lda;
  const float* B;
  int ldb;
  float* C;
  int ldc;
  const float* aux;      // bias [N] (T_BIAS*), or the layer input [M, ldaux] (T_RELU_MASK)
  int ldaux;
  float* db;             // MODE_TN: column sums of P (nullptr: not wanted)
  int M;                 // rows (multiple of 128)
  int K;                 // MODE_NT: contraction length
  int n_valid;           // MODE_NT: output columns (<= 256);  MODE_TN: columns of Q
  int p_valid;           // MODE_TN: columns of P
  int splits;            // MODE_TN: CTAs along the rows
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
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void split_store4(uint8_t* big, uint8_t* small, const float4& v) {
  float4 b, s;
  b.x = tf32_rna(v.x); b.y = tf32_rna(v.y); b.z = tf32_rna(v.z); b.w = tf32_rna(v.w);
  s.x = tf32_rna(v.x - b.x); s.y = tf32_rna(v.y - b.y); s.z = tf32_rna(v.z - b.z); s.w = tf32_rna(v.w - b.w);
  *reinterpret_cast<float4*>(big) = b;
  *reinterpret_cast<float4*>(small) = s;
}

// out[c][r] = in[r][c]  (the dX layers read W^T as a K-major operand)
static __global__ void k_tf32_transpose(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out, int ld_out) {
  pdl_wait();
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * ld_out + r] = t[threadIdx.x][i];
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
  uint64_t* full = bars;                    // [kStages]  8 loader-warp arrivals
  uint64_t* empty = bars + kStages;         // [kStages]  MMAs of the stage retired
  uint64_t* acc_full = bars + 2 * kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work of this CTA
  //   MODE_NT: 128-row tiles bid, bid + grid, ... each over ceil(K / 32) stages
  //   MODE_TN: one accumulator; a contiguous range of the 32-row stages
  int n_tiles_my, st_begin = 0, st_count;
  if (MODE == MODE_NT) {
    const int n_tiles = p.M / kTileM;
    n_tiles_my = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    st_count = (p.K + kStageK - 1) / kStageK;
  } else {
    const int st_total = p.M / kStageK;
    const int per = (st_total + p.splits - 1) / p.splits;
    st_begin = (int)blockIdx.x * per;
    st_count = max(0, min(per, st_total - st_begin));
    n_tiles_my = st_count > 0 ? 1 : 0;
  }
  const int p0 = MODE == MODE_TN ? (int)blockIdx.y * 128 : 0;     // first column of P (= first row of C) of this CTA
  const int q0 = MODE == MODE_TN ? (int)blockIdx.z * 256 : 0;     // first column of Q (= first column of C)
  const int nv = MODE == MODE_TN ? min(256, p.n_valid - q0) : p.n_valid;      // valid accumulator columns
  const int npad = (nv + 15) / 16 * 16;                                       // MMA N
  const int pv = MODE == MODE_TN ? min(128, p.p_valid - p0) : 128;            // valid accumulator rows

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                // everything below reads the previous kernels' output

  if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(kTileM, npad);
      uint32_t it = 0;
      for (int t = 0; t < n_tiles_my; ++t) {
        const uint32_t a = t & 1, aph = (t >> 1) & 1;
        mbar_wait(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * 256;
        for (int ks = 0; ks < st_count; ++ks, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (p.trace && blockIdx.x == 0 && it < 512) p.trace[it * 8 + 4] = clock64();
          const uint32_t sa = smem_u32(smem + s * kStageBytes);
          const uint64_t da_big = smem_desc_sw128(sa, 16, 1024);
          const uint64_t da_small = smem_desc_sw128(sa + kPlaneA, 16, 1024);
          const uint64_t db_big = smem_desc_sw128(sa + 2 * kPlaneA, 16, 1024);
          const uint64_t db_small = smem_desc_sw128(sa + 2 * kPlaneA + kPlaneB, 16, 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) {       // 4 K-steps of 8 tf32 (32 bytes) per stage; small terms first
            umma_tf32(d_tmem, da_small + 2 * j, db_big + 2 * j, idesc, (ks | j) != 0);
            umma_tf32(d_tmem, da_big + 2 * j, db_small + 2 * j, idesc, 1);
            umma_tf32(d_tmem, da_big + 2 * j, db_big + 2 * j, idesc, 1);
          }
          umma_commit(&empty[s]);
          if (p.trace && blockIdx.x == 0 && it < 512) p.trace[it * 8 + 5] = clock64();
        }
        umma_commit(&acc_full[a]);
      }
    }
    __syncwarp();
  }
  if (warp >= kLoaderWarp0 && warp < kEpiWarp0) {
    // ------------------------------------------------------------------ loaders: split (and transpose) into the stages
    const int lt = threadIdx.x - 32 * kLoaderWarp0;          // 0..255
    const int lw = lt >> 5;                   // loader warp 0..7
    // two register sets = the operands of two stages in flight: the loads of stage i + 2 are issued as soon as the registers
    // of stage i have been written to SMEM (they do not wait for an SMEM slot), so the load latency is hidden behind two
    // stages of MMA time
    float4 va[12], vb[12];
    const int n_it = n_tiles_my * st_count;   // stages of this CTA (MODE_TN: n_tiles_my = 1)
    if (MODE == MODE_NT) {
      // chunk (row0 + 32 j, kc): row0 = lt >> 3 (0..31), kc = lt & 7; j = 0..3 for A (128 rows), 0..7 for B (256 rows);
      // 8 lanes read one 128-byte row segment, the SW128 offset of (row0 + 32 j, kc) is base + 4096 j
      const int row0 = lt >> 3, kc = lt & 7;
      const uint32_t base = (uint32_t)(row0 >> 3) * 1024u + (uint32_t)(row0 & 7) * 128u + (uint32_t)((kc ^ (row0 & 7)) << 4);
      const float* gB = p.B + (size_t)row0 * p.ldb + kc * 4;
      int nB = 0;                             // B chunks of this thread that lie inside the matrix (the rest is zero)
#pragma unroll
      for (int j = 0; j < 8; ++j) nB += row0 + 32 * j < nv ? 1 : 0;
      const int nBs = (npad - row0 + 31) / 32;          // chunks that the MMA reads (rows < npad)
      auto issue = [&](float4 (&v)[12], int i) {
        if (i >= n_it) return;
        const int t = i / st_count, ks = i - t * st_count;
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int k = ks * kStageK + kc * 4;
        const bool kin = k < p.K;
        const float* gA = p.A + (size_t)(tile * kTileM + row0) * p.lda + k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kin) v[j] = *reinterpret_cast<const float4*>(gA + (size_t)(32 * j) * p.lda);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kin && j < nB) v[4 + j] = *reinterpret_cast<const float4*>(gB + (size_t)(32 * j) * p.ldb + ks * kStageK);
        }
      };
      auto stage = [&](float4 (&v)[12], int i) {
        if (i >= n_it) return;
        const uint32_t s = (uint32_t)i % kStages, ph = ((uint32_t)i / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const bool tr = p.trace && blockIdx.x == 0 && lt == 0 && i < 512;
        if (tr) p.trace[i * 8 + 0] = clock64();
        uint8_t* sa = smem + s * kStageBytes + base;
#pragma unroll
        for (int j = 0; j < 4; ++j) split_store4(sa + j * 4096, sa + kPlaneA + j * 4096, v[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < nBs) split_store4(sa + 2 * kPlaneA + j * 4096, sa + 2 * kPlaneA + kPlaneB + j * 4096, v[4 + j]);
        if (tr) p.trace[i * 8 + 1] = clock64();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (tr) p.trace[i * 8 + 2] = clock64();
        issue(v, i + 2);
      };
      issue(va, 0);
      issue(vb, 1);
      for (int i = 0; i < n_it; i += 2) {
        stage(va, i);
        stage(vb, i + 1);
      }
    } else {
      // thread -> row m of the stage and column chunk c4_0 (of 4 floats); unit j adds 8 chunks (32 columns): P for j < 4
      // (128 columns), Q for j = 4..11 (256 columns).  Component i of a chunk goes to operand row 4 c4 + i, element m:
      // SW128 offset = base[i] + 4096 j.  Within a warp m covers 16 values and c4 two (lane >> 4): rows 4 c4 + i of the two
      // halves differ by 4 in their swizzle phase, so the 32 lanes hit 32 distinct banks; the loads are 16 x 32-byte sectors.
      const int m = (lw & 1) * 16 + (lane & 15);
      const int c4_0 = (lw >> 1) * 2 + (lane >> 4);      // 0..7
      uint32_t base[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) base[i] = sw128_off((uint32_t)(4 * c4_0 + i), (uint32_t)m);
      const float* gP = p.A + (size_t)m * p.lda + p0 + 4 * c4_0;
      const float* gQ = p.B + (size_t)m * p.ldb + q0 + 4 * c4_0;
      int nP = 0, nQ = 0;                      // units inside the matrices
#pragma unroll
      for (int j = 0; j < 4; ++j) nP += 4 * c4_0 + 32 * j < pv ? 1 : 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) nQ += 4 * c4_0 + 32 * j < nv ? 1 : 0;
      const int nQs = (npad - 4 * c4_0 + 31) / 32;       // units that the MMA reads (operand rows < npad)
      float colsum[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) colsum[i] = 0.f;
      auto issue = [&](float4 (&v)[12], int i) {
        if (i >= n_it) return;
        const size_t r = (size_t)(st_begin + i) * kStageK;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nP) v[j] = *reinterpret_cast<const float4*>(gP + r * p.lda + 32 * j);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < nQ) v[4 + j] = *reinterpret_cast<const float4*>(gQ + r * p.ldb + 32 * j);
        }
      };
      auto stage = [&](float4 (&v)[12], int i) {
        if (i >= n_it) return;
        const uint32_t s = (uint32_t)i % kStages, ph = ((uint32_t)i / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && lt == 0 && i < 512;
        if (tr) p.trace[i * 8 + 0] = clock64();
        uint8_t* sa = smem + s * kStageBytes;
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          if (j >= 4 && j - 4 >= nQs) continue;
          uint8_t* pl = sa + (j < 4 ? j * 4096 : 2 * kPlaneA + (j - 4) * 4096);
          const uint32_t pstride = j < 4 ? kPlaneA : kPlaneB;
          const float x[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float b = tf32_rna(x[c]);
            *reinterpret_cast<float*>(pl + base[c]) = b;
            *reinterpret_cast<float*>(pl + pstride + base[c]) = tf32_rna(x[c] - b);
            if (j < 4) colsum[j * 4 + c] += x[c];
          }
        }
        if (tr) p.trace[i * 8 + 1] = clock64();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
        if (tr) p.trace[i * 8 + 2] = clock64();
        issue(v, i + 2);
      };
      issue(va, 0);
      issue(vb, 1);
      for (int i = 0; i < n_it; i += 2) {
        stage(va, i);
        stage(vb, i + 1);
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
    // ------------------------------------------------------------------ epilogue (4 warps)
    const int q = warp & 3;
    uint8_t* stg = sOut + q * 4096;            // this warp's [32 rows x 32 columns] staging block, 16-byte chunks swizzled by row
    if ((EPI == T_BIAS || EPI == T_BIAS_RELU) && MODE == MODE_NT) {
      for (int i = threadIdx.x - 32 * kEpiWarp0; i < 256; i += 128) sBias[i] = i < nv ? p.aux[i] : 0.f;
      named_bar_sync(1, 128);
    }
    const int n_groups = (npad + 31) / 32;
    const int rr = lane >> 3, ch = lane & 7;   // read-back: rows rr + 4 i, chunk ch -> 8 lanes cover one 128-byte line
    // T_RELU_MASK: the mask (the layer input) does not depend on the accumulator: the 8 lines a thread needs for a column
    // group are requested one group ahead (the first ones before the accumulator is even complete)
    float4 mk[8];
    auto load_mask = [&](int t, int g) {
      if (EPI != T_RELU_MASK || MODE != MODE_NT || t >= n_tiles_my) return;
      const int col = g * 32 + ch * 4;
      const size_t grow0 = (size_t)((int)blockIdx.x + t * (int)gridDim.x) * kTileM + q * 32 + rr;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        mk[i] = col < nv ? *reinterpret_cast<const float4*>(p.aux + (grow0 + 4 * i) * p.ldaux + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    load_mask(0, 0);
    for (int t = 0; t < n_tiles_my; ++t) {
      const uint32_t a = t & 1, aph = (t >> 1) & 1;
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      const bool tre = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 32 * kEpiWarp0 && t < 64;
      if (tre) p.trace[t * 8 + 6] = clock64();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + a * 256;
      uint32_t v[32];
      tmem_ld32(tbase, v);
      for (int g = 0; g < n_groups; ++g) {
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = make_uint4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        if (g + 1 < n_groups) tmem_ld32(tbase + (g + 1) * 32, v);       // next group's accumulator columns travel meanwhile
        __syncwarp();
        uint32_t curbits = 0;                    // bit 4 i + c: element c of row rr + 4 i passes
        if (EPI == T_RELU_MASK) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            curbits |= ((mk[i].x > 0.f ? 1u : 0u) | (mk[i].y > 0.f ? 2u : 0u) | (mk[i].z > 0.f ? 4u : 0u) | (mk[i].w > 0.f ? 8u : 0u)) << (4 * i);
        }
        if (g + 1 < n_groups) load_mask(t, g + 1); else load_mask(t + 1, 0);
        const int col = g * 32 + ch * 4;
        if (col < nv) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = rr + 4 * i;          // row inside this warp's 32
            float4 y = *reinterpret_cast<const float4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
            if (MODE == MODE_NT) {
              const size_t grow = (size_t)((int)blockIdx.x + t * (int)gridDim.x) * kTileM + q * 32 + row;
              if (EPI == T_BIAS || EPI == T_BIAS_RELU) {
                const float4 b = *reinterpret_cast<const float4*>(sBias + col);
                y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
                if (EPI == T_BIAS_RELU) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
              } else if (EPI == T_RELU_MASK) {
                const uint32_t b4 = curbits >> (4 * i);
                y.x = (b4 & 1u) ? y.x : 0.f; y.y = (b4 & 2u) ? y.y : 0.f; y.z = (b4 & 4u) ? y.z : 0.f; y.w = (b4 & 8u) ? y.w : 0.f;
              }
              *reinterpret_cast<float4*>(p.C + grow * p.ldc + col) = y;
            } else {
              const int crow = q * 32 + row;
              if (crow < pv) red_add_v4(p.C + (size_t)(p0 + crow) * p.ldc + q0 + col, y.x, y.y, y.z, y.w);
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
      if (tre) p.trace[t * 8 + 7] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace t32
}  // namespace marf
