// tcgen05 kernels of the bf16 tensor-core path (sm_100a).
//
//   k_tc_gemm<N_TILE,EPI> : C[128-row tiles, N_TILE] = epi(A[rows,K] * W[N,K]^T), A streamed by TMA, W resident in SMEM,
//                           fp32 accumulators double-buffered in TMEM, epilogue through SMEM + TMA store.
//                           Used for the forward layers (bias+ReLU) and the dX layers (ReLU mask of the layer input).
//   k_tc_dw<N_TILE>       : dW[out,in] += dY[rows,out]^T * X[rows,in] (contraction over pixel rows, both operands
//                           MN-major straight from the row-major activations), split over CTAs, fp32 red.add at the end.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
// (warp w owns TMEM lanes 32*(w%4)..+31 = tile rows).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace marf {
namespace tc {

constexpr int kTileM = 128;          // pixel rows per tile (= TMEM lanes)
constexpr int kChunkK = 64;          // bf16 elements per 128-byte swizzle row
constexpr int kChunkBytes = kTileM * 128;   // one [128 x 64] bf16 operand chunk
constexpr int kStages = 4;
constexpr int kThreads = 320;         // k_tc_gemm: producer warp, MMA warp, 8 epilogue warps
constexpr int kDwThreads = 192;       // k_tc_dw: producer warp, MMA warp, 4 epilogue warps

enum { EPI_BIAS_RELU = 0, EPI_RELU_MASK = 1, EPI_PLAIN_F32 = 2, EPI_WARP_GRAD = 3 };

struct GemmParams {
  int n_tiles;          // 128-row tiles
  int k_chunks;         // K / 64
  const float* bias;    // [N] (EPI_BIAS_RELU)
  float* out_f32;       // EPI_PLAIN_F32: [rows, ld_out] fp32, first n_store columns written
  int ld_out;
  int n_store;
  uint32_t* bits_out;   // EPI_BIAS_RELU: optional [rows, bits_ld] words, bit c of a row = (output column c > 0)
  const uint32_t* bits_in;  // EPI_RELU_MASK: the ReLU mask of the layer input, same layout
  int bits_ld;          // words per row (= padded width / 32)
  int reverse;          // walk the tiles last-to-first (the previous layer's newest output is still in L2)
  unsigned long long load_policy;   // L2 cache hint of the streamed A tiles
  unsigned long long store_policy;  // L2 cache hint of the output tiles
  int prefetch_ahead;   // tiles to prefetch into L2 ahead of the SMEM ring (0 = off)
  // EPI_WARP_GRAD (N_TILE = 64): the tile is dL/d(encoded input); the epilogue runs the backward of the posenc + homography
  // prologue per pixel and reduces the per-patch 3x3 Jacobian G (fp64 atomics), nothing is stored per pixel
  Geo geo;
  PxRange rg;
  const float* Hm;      // [batch_global, 9]
  double* G;            // [batch, 9]
  long long* trace;     // diagnostics: per-tile clock64() stamps of CTA 0 ([tile_iter][16]); nullptr in production
};

struct GemmSmem {       // offsets computed on host and device identically
  int w_bytes, a_off, out_off, bias_off, bar_off, total;
};
__host__ __device__ inline GemmSmem gemm_smem(int n_tile, int k_chunks, bool staged_out) {
  GemmSmem s;
  s.w_bytes = k_chunks * n_tile * 128;
  s.a_off = s.w_bytes;
  s.out_off = s.a_off + kStages * kChunkBytes;
  s.bias_off = s.out_off + (staged_out ? 2 * kChunkBytes : 0);
  s.bar_off = s.bias_off + n_tile * 4;
  s.total = s.bar_off + 256;
  return s;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// One launch = several GEMM "jobs" (layers); a contiguous range of CTAs serves each job, persistent over its tiles.
// Jobs may be chained: job B whose input is job A's output waits, tile by tile, on A's completion flags, so the
// activations travel producer -> consumer through L2 instead of HBM (all CTAs of the launch are co-resident:
// grid <= #SMs, 1 CTA/SM).
struct GemmJob {
  CUtensorMap tmA, tmW, tmOut;
  GemmParams p;
  int n0;                      // first output column of this job (N tiles of one layer are separate jobs)
  int cta_begin, cta_count;    // CTAs [cta_begin, cta_begin+cta_count) of the launch work on this job
  const uint32_t* flags_in;    // per-tile completion counters of the producer job (nullptr: input already complete)
  uint32_t* flags_out;         // per-tile completion counters this job bumps (nullptr: nobody waits)
  uint32_t in_target;          // flags_in[tile] >= in_target  <=>  the input tile is complete
};
constexpr int kMaxGemmJobs = 8;
struct GemmJobs { GemmJob j[kMaxGemmJobs]; int n; };

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int N_TILE, int EPI, int LT = 0>   // LT: compile-time posenc band count for EPI_WARP_GRAD (0: runtime)
__global__ void __launch_bounds__(kThreads, 1) k_tc_gemm(const __grid_constant__ GemmJobs jobs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr bool kStaged = EPI != EPI_PLAIN_F32 && EPI != EPI_WARP_GRAD;
  constexpr int kSlabs = N_TILE / 64;
  constexpr uint32_t kTmemCols = 2 * N_TILE < 32 ? 32 : 2 * N_TILE;
  int jr = 0;
  for (int i = 1; i < jobs.n; ++i)
    if ((int)blockIdx.x >= jobs.j[i].cta_begin) jr = i;
  const GemmJob& J = jobs.j[jr];
  const GemmParams& p = J.p;
  const int bid = (int)blockIdx.x - J.cta_begin;
  const int nctas = J.cta_count;
  if (bid >= nctas) return;                          // (grid padded: not used)
  const GemmSmem L = gemm_smem(N_TILE, p.k_chunks, kStaged);
  uint8_t* sW = smem;
  uint8_t* sA = smem + L.a_off;
  uint8_t* sOut = smem + L.out_off;
  float* sBias = reinterpret_cast<float*>(smem + L.bias_off);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* w_full = bars + 2 * kStages;
  uint64_t* acc_full = w_full + 1;       // [2]
  uint64_t* acc_empty = acc_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = J.n0;
  // tiles of this CTA: strided (bid, bid + nctas, ...) by default; EPI_WARP_GRAD walks a contiguous block of tiles per CTA so
  // that a CTA stays inside one or two patches and its per-patch partial sums stay in registers
  const int blk_per = (p.n_tiles + nctas - 1) / nctas;
  const int n_my = EPI == EPI_WARP_GRAD ? max(0, min(blk_per, p.n_tiles - bid * blk_per))
                                        : (bid < p.n_tiles ? (p.n_tiles - bid + nctas - 1) / nctas : 0);
  auto tile_at = [&](int i) -> int {
    if (EPI == EPI_WARP_GRAD) return bid * blk_per + i;
    const int t0 = bid + i * nctas;
    return p.reverse ? p.n_tiles - 1 - t0 : t0;
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&J.tmA);
    prefetch_tmap(&J.tmW);
    if (kStaged) prefetch_tmap(&J.tmOut);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(w_full, 1);
    // (EPI_WARP_GRAD: the two epilogue groups take alternate tiles = alternate accumulators, 4 warps release each)
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI == EPI_WARP_GRAD ? 4 : 8); }
    fence_barrier_init();
    // the weights were packed at least two launches ago: fetch them while the previous kernel is still draining
    mbar_expect_tx(w_full, (uint32_t)L.w_bytes);
    for (int c = 0; c < p.k_chunks; ++c) tma_load_2d(sW + c * N_TILE * 128, &J.tmW, c * kChunkK, n0, w_full);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  if (EPI == EPI_BIAS_RELU && warp >= 2)
    for (int i = threadIdx.x - 64; i < N_TILE; i += 256) sBias[i] = p.bias[n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: everything below consumes the previous kernel's output

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      const int kAhead = p.prefetch_ahead;           // tiles prefetched into L2 ahead of the SMEM ring (0: off)
      if (!J.flags_in && kAhead > 0)
        for (int a = 0; a < kAhead; ++a) {
          const int tp = bid + a * nctas;
          if (tp < p.n_tiles)
            for (int c = 0; c < p.k_chunks; ++c) tma_prefetch_2d(&J.tmA, c * kChunkK, (p.reverse ? p.n_tiles - 1 - tp : tp) * kTileM);
        }
      for (int i = 0; i < n_my; ++i) {
        const int t0 = bid + i * nctas;
        const int tile = tile_at(i);
        if (!J.flags_in && kAhead > 0) {
          const int tp = t0 + kAhead * nctas;
          if (tp < p.n_tiles)
            for (int c = 0; c < p.k_chunks; ++c) tma_prefetch_2d(&J.tmA, c * kChunkK, (p.reverse ? p.n_tiles - 1 - tp : tp) * kTileM);
        }
        if (J.flags_in) {
          // wait until the producer job has finished this tile (bounded: a protocol bug traps instead of hanging)
          uint32_t spins = 0;
          while (ld_acquire_gpu(J.flags_in + tile) < J.in_target) {
            __nanosleep(64);
            if (++spins > (1u << 22)) __trap();
          }
          fence_proxy_async_all();                   // order the acquire before the async-proxy (TMA) reads
        }
        for (int c = 0; c < p.k_chunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], kChunkBytes);
          tma_load_2d_hint(sA + s * kChunkBytes, &J.tmA, c * kChunkK, tile * kTileM, &full[s], p.load_policy);
        }
        if (p.trace && bid == 0) p.trace[i * 16 + 0] = clock64();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(kTileM, N_TILE, 0, 0);
      mbar_wait(w_full, 0);
      tc_fence_after();
      uint32_t it = 0, t_iter = 0;
      for (int i = 0; i < n_my; ++i, ++t_iter) {
        const uint32_t a = t_iter & 1, aph = (t_iter >> 1) & 1;
        mbar_wait(&acc_empty[a], aph ^ 1);
        tc_fence_after();
        if (p.trace && bid == 0) p.trace[t_iter * 16 + 1] = clock64();
        const uint32_t d_tmem = tmem_base + a * N_TILE;
        for (int c = 0; c < p.k_chunks; ++c, ++it) {
          const uint32_t s = it % kStages, ph = (it / kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint64_t da = smem_desc_sw128(smem_u32(sA + s * kChunkBytes), 16, 1024);
          const uint64_t db = smem_desc_sw128(smem_u32(sW + c * N_TILE * 128), 16, 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(d_tmem, da + 2 * j, db + 2 * j, idesc, (c | j) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[a]);
        if (p.trace && bid == 0) p.trace[t_iter * 16 + 2] = clock64();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps (256 threads)
    // group g handles the 64-column slabs g, g+2, ...; each group owns one staging buffer and two named barriers.
    const int q = warp & 3;                            // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;                       // tile row == TMEM lane
    const bool gleader = threadIdx.x == 64 + 128 * grp;
    const uint32_t sw_row = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    uint8_t* ob = sOut + grp * kChunkBytes;
    constexpr int kLastSlabOfGroup0 = ((kSlabs - 1) & 1) == 0 ? kSlabs - 1 : kSlabs - 2;
    constexpr int kLastSlabOfGroup1 = ((kSlabs - 1) & 1) == 1 ? kSlabs - 1 : kSlabs - 2;
    constexpr int kGroupSlabs = kSlabs >= 2 ? kSlabs / 2 : 1;      // stores a group commits per tile
    // flags_out[tile] receives +1 from every group that stores slabs of the tile (2 groups when kSlabs >= 2)
    int pending_tile = -1;                             // tile whose stores are committed but not yet signalled
    uint32_t t_iter = 0;
    // EPI_WARP_GRAD: per-warp partial sums of the current patch's 3x3 Jacobian, flushed (one fp64 atomic per entry) when the
    // patch changes and at the end: the fp64 atomics on the few [batch, 9] addresses serialise in L2 otherwise
    float wg_acc[9];
    int wg_patch = -1;
#pragma unroll
    for (int i = 0; i < 9; ++i) wg_acc[i] = 0.f;
    for (int i_t = 0; i_t < n_my; ++i_t, ++t_iter) {
      if (EPI == EPI_WARP_GRAD && (int)(t_iter & 1) != grp) continue;      // the other group's tile
      const int tile = tile_at(i_t);
      const uint32_t a = t_iter & 1, aph = (t_iter >> 1) & 1;
      uint32_t mbits[2 * kSlabs];
      if (EPI == EPI_RELU_MASK) {
        const uint32_t* bp = p.bits_in + (size_t)(tile * kTileM + r) * p.bits_ld + n0 / 32;
#pragma unroll
        for (int i = 0; i < 2 * kSlabs; ++i) mbits[i] = bp[i];
      }
      mbar_wait(&acc_full[a], aph);
      tc_fence_after();
      const bool tr = p.trace && bid == 0 && (threadIdx.x == 64 || threadIdx.x == 192);
      if (tr) p.trace[t_iter * 16 + 3 + grp * 6] = clock64();
#pragma unroll
      for (int j = 0; j < kSlabs; ++j) {
        if (EPI != EPI_WARP_GRAD && (j & 1) != grp) continue;
        uint32_t v[64];
        {
          uint32_t (&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          uint32_t (&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + a * N_TILE + j * 64;
          tmem_ld32(taddr, v0);
          tmem_ld32(taddr + 32, v1);
        }
        if (kStaged) {
          if (gleader) bulk_wait_read<0>();            // this group's previous store has finished reading the buffer
          named_bar_sync(1 + 2 * grp, 128);
        }
        tmem_ld_wait();
        if (tr) p.trace[t_iter * 16 + 4 + grp * 6 + (j >> 1) * 2] = clock64();
        if (EPI == EPI_WARP_GRAD) {
          const int t = tile * kTileM + r;
          const bool valid = t < p.rg.count;
          int b = -1;
          float acc[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) acc[i] = 0.f;
          if (valid) {
            int rr, cc;
            decode_px(p.geo, p.rg.first + t, b, rr, cc);
            float x, y, u, vv, qz;
            grid_xy(p.geo, rr, cc, x, y);
            apply_h(p.Hm + 9 * (b + p.geo.patch_offset), x, y, u, vv, qz);
            const bool do_u = true, do_v = true;
            float gu = do_u ? __uint_as_float(v[0]) : 0.f, gv = do_v ? __uint_as_float(v[1]) : 0.f;
            const int Lb = LT > 0 ? LT : p.geo.L;
#pragma unroll
            for (int k = 0; k < (LT > 0 ? LT : 15); ++k) {
              if (k < Lb) {
                float su = 0.f, cu = 0.f, sv = 0.f, cv = 0.f;
                const float f = p.geo.band_f[k];
                if (do_u) sincosf(u * f, &su, &cu);
                if (do_v) sincosf(vv * f, &sv, &cv);
                const float wf = p.geo.band_w[k] * f;
                float d_su = 0.f, d_cu = 0.f, d_sv = 0.f, d_cv = 0.f;
                if (LT > 0) {                          // static register indices
                  d_su = __uint_as_float(v[(2 + k) & 63]);
                  d_cu = __uint_as_float(v[(2 + LT + k) & 63]);
                  d_sv = __uint_as_float(v[(2 + 2 * LT + k) & 63]);
                  d_cv = __uint_as_float(v[(2 + 3 * LT + k) & 63]);
                } else {                               // runtime L: select with an unrolled scan
#pragma unroll
                  for (int c2 = 2; c2 < 62; ++c2) {
                    const float val = __uint_as_float(v[c2]);
                    if (c2 == 2 + k) d_su = val;
                    if (c2 == 2 + Lb + k) d_cu = val;
                    if (c2 == 2 + 2 * Lb + k) d_sv = val;
                    if (c2 == 2 + 3 * Lb + k) d_cv = val;
                  }
                }
                gu += wf * (d_su * cu - d_cu * su);
                gv += wf * (d_sv * cv - d_cv * sv);
              }
            }
            const float inv = 1.0f / qz;
            const float dq0 = gu * inv, dq1 = gv * inv, dq2 = -(gu * u + gv * vv) * inv;
            acc[0] = dq0 * x; acc[1] = dq0 * y; acc[2] = dq0;
            acc[3] = dq1 * x; acc[4] = dq1 * y; acc[5] = dq1;
            acc[6] = dq2 * x; acc[7] = dq2 * y; acc[8] = dq2;
          }
          const int b0 = __shfl_sync(0xffffffffu, b, 0);
          const bool uniform = __all_sync(0xffffffffu, b == b0 || b < 0);
          if (uniform) {
            if (b0 >= 0) {
              if (b0 != wg_patch) {
                if (wg_patch >= 0 && lane == 0) {
#pragma unroll
                  for (int i = 0; i < 9; ++i) atomicAdd(&p.G[9 * wg_patch + i], (double)wg_acc[i]);
                }
#pragma unroll
                for (int i = 0; i < 9; ++i) wg_acc[i] = 0.f;
                wg_patch = b0;
              }
#pragma unroll
              for (int i = 0; i < 9; ++i) wg_acc[i] += warp_sum(acc[i]);      // (every lane keeps the same running sum)
            }
          } else if (valid) {
#pragma unroll
            for (int i = 0; i < 9; ++i) atomicAdd(&p.G[9 * b + i], (double)acc[i]);
          }
        } else if (EPI == EPI_PLAIN_F32) {
          float* o = p.out_f32 + (size_t)(tile * kTileM + r) * p.ld_out + n0 + j * 64;
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4)
            if (n0 + j * 64 + c4 * 4 < p.n_store)
              *reinterpret_cast<uint4*>(o + c4 * 4) = make_uint4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        } else {
          // mask-bit layout inside a 32-column group: bit t <- column 2t, bit 16+t <- column 2t+1 (t = 0..15)
          uint32_t obits[2] = {0u, 0u};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint32_t w[4];
            float bv[8];
            if (EPI == EPI_BIAS_RELU) {
              const float4 b0 = *reinterpret_cast<const float4*>(sBias + j * 64 + i * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sBias + j * 64 + i * 8 + 4);
              bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
            }
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
              const int c = i * 8 + pr * 2;             // column inside the slab
              const int t = (c & 31) >> 1;              // pair index inside the 32-column group
              float lo = __uint_as_float(v[c]), hi = __uint_as_float(v[c + 1]);
              if (EPI == EPI_BIAS_RELU) {
                lo += bv[pr * 2];
                hi += bv[pr * 2 + 1];
                asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                uint32_t gt;
                asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(gt) : "r"(w[pr]), "r"(0u));
                obits[c >> 5] |= gt & (0x00010001u << t);
              } else {
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                const uint32_t sel = (mbits[2 * j + (c >> 5)] >> t) & 0x00010001u;
                w[pr] &= sel * 0xFFFFu;
              }
            }
            *reinterpret_cast<uint4*>(ob + sw_row + (((uint32_t)i ^ (uint32_t)(r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if (EPI == EPI_BIAS_RELU && p.bits_out)
            *reinterpret_cast<uint2*>(p.bits_out + (size_t)(tile * kTileM + r) * p.bits_ld + (n0 + j * 64) / 32) =
                make_uint2(obits[0], obits[1]);
          fence_proxy_async_smem();
          named_bar_sync(2 + 2 * grp, 128);
          if (tr) p.trace[t_iter * 16 + 5 + grp * 6 + (j >> 1) * 2] = clock64();
          if (gleader) {
            tma_store_2d_hint(&J.tmOut, n0 + j * 64, tile * kTileM, ob, p.store_policy);
            bulk_commit();
            if (J.flags_out && j == (grp == 0 ? kLastSlabOfGroup0 : kLastSlabOfGroup1)) {
              // lagged signalling: a TMA store needs ~3k cycles to complete, so the tile finished ONE tile ago is
              // published now: everything except this tile's own stores (kGroupSlabs newest groups) has completed
              if (pending_tile >= 0) {
                bulk_wait<kGroupSlabs>();
                fence_proxy_async_all();
                red_release_gpu_add(J.flags_out + pending_tile, 1u);
              }
              pending_tile = tile;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);
    }
    if (EPI == EPI_WARP_GRAD && wg_patch >= 0 && lane == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) atomicAdd(&p.G[9 * wg_patch + i], (double)wg_acc[i]);
    }
    if (kStaged && gleader) {
      bulk_wait<0>();
      if (J.flags_out && pending_tile >= 0) {
        fence_proxy_async_all();
        red_release_gpu_add(J.flags_out + pending_tile, 1u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// ================================================================================================
// dW[out, in] += sum over pixel rows of dY[row, out] * X[row, in];  db[out] += sum over rows of dY[row, out]
// One launch covers several layers ("jobs", blockIdx.y); blockIdx.x splits the pixel rows of a job.
// While the MMA warp streams the stages, the four epilogue warps read the dY stage from SMEM and keep the
// bias-gradient column sums in registers (no extra pass over dY in HBM).
// ================================================================================================
struct DwJob {
  CUtensorMap tmDY;     // box {64,64} over dY [rows, out]
  CUtensorMap tmX;      // box {64,64} over X  [rows, in]
  int rows;             // padded pixel rows (multiple of 64)
  int rows_per_cta;     // multiple of 64
  int cta_begin, cta_count;   // CTAs [cta_begin, cta_begin + cta_count) of the launch split this job's pixel rows
  int n_tile;           // input columns of this job's tile: 64 or 256 (the MMA N)
  int m0;               // first output feature of this job (outputs wider than 256 are split over jobs)
  int m_halves;         // 128-row halves of dY^T in this job (1 or 2)
  int m_valid;          // out features of the layer
  int n_valid;          // in features
  int n0;               // first input column of this job's N tile
  int ld_w;
  int do_bias;          // accumulate db (only one N tile per layer does)
  float* dW;            // [out, ld_w] fp32, accumulated with red.add
  float* db;            // [out] fp32
};
constexpr int kDwMaxJobs = 12;
struct DwJobs { DwJob j[kDwMaxJobs]; int n; };

constexpr int kDwStages = 3;
constexpr int kDwRows = 64;                 // pixel rows per stage
constexpr int kDwSlab = kDwRows * 128;      // [64 rows x 64 cols] bf16

// One launch covers every layer of both networks ("jobs"); a contiguous range of CTAs splits the pixel rows of a job
// (ranges sized by the bytes the job streams).  The tile width N (64 for the encoded-input layers, 256 otherwise) is a
// per-job runtime value: it only enters the instruction descriptor, the stage layout and the epilogue bounds.
__global__ void __launch_bounds__(kDwThreads, 1) k_tc_dw(const __grid_constant__ DwJobs jobs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  int jr = 0;
  for (int i = 1; i < jobs.n; ++i)
    if ((int)blockIdx.x >= jobs.j[i].cta_begin) jr = i;
  const DwJob& J = jobs.j[jr];
  const int bid = (int)blockIdx.x - J.cta_begin;
  const int n_tile = J.n_tile;
  const int b_slabs = n_tile / 64;
  const int a_slabs = J.m_halves * 2;
  const int stage_bytes = (a_slabs + b_slabs) * kDwSlab;
  const int stage_stride = (4 + 4) * kDwSlab;                 // stages are laid out for the largest job
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * stage_stride);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDwStages;
  uint64_t* done = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const uint32_t tmem_cols = 512;   // m_halves * n_tile <= 512

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_begin = bid * J.rows_per_cta;
  const int row_end = min(J.rows, row_begin + J.rows_per_cta);
  const int n_iter = (bid < J.cta_count && row_begin < row_end) ? (row_end - row_begin) / kDwRows : 0;

  if (threadIdx.x == 0) {
    prefetch_tmap(&J.tmDY);
    prefetch_tmap(&J.tmX);
    for (int s = 0; s < kDwStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 5); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL

  if (n_iter > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
          uint8_t* st = smem + s * stage_stride;
          const int row = row_begin + it * kDwRows;
          for (int i = 0; i < a_slabs; ++i) tma_load_2d(st + i * kDwSlab, &J.tmDY, J.m0 + i * 64, row, &full[s]);
          for (int i = 0; i < b_slabs; ++i) tma_load_2d(st + (a_slabs + i) * kDwSlab, &J.tmX, J.n0 + i * 64, row, &full[s]);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = idesc_bf16(128, n_tile, 1, 1);     // both operands MN-major
        for (int it = 0; it < n_iter; ++it) {
          const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * stage_stride);
#pragma unroll
          for (int ks = 0; ks < kDwRows / 16; ++ks) {
            const uint64_t db = smem_desc_sw128(st + a_slabs * kDwSlab + ks * 2048, kDwSlab, 1024);
            // Both 128-row halves of dY^T unconditionally: for a one-half job (output layers) the second MMA multiplies
            // whatever lies behind the two dY slabs into TMEM columns nobody reads.  A branch here would be if-converted
            // by ptxas into a predicated UTCHMMA (see tc_chain.cuh) — every tcgen05.mma stays unconditional.
            const uint64_t da0 = smem_desc_sw128(st + ks * 2048, kDwSlab, 1024);
            const uint64_t da1 = smem_desc_sw128(st + 2 * kDwSlab + ks * 2048, kDwSlab, 1024);
            umma_bf16(tmem_base, da0, db, idesc, (it | ks) != 0);
            umma_bf16(tmem_base + n_tile, da1, db, idesc, (it | ks) != 0);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(done);
      }
    } else {
      const int q = warp & 3;
      const int r = q * 32 + lane;
      // ---- bias gradient: warp q sums the 64 out-columns of dY slab q (if present) over the stage's 64 rows
      float bs0 = 0.f, bs1 = 0.f;
      const bool bias_warp = J.do_bias && q < a_slabs;
      for (int it = 0; it < n_iter; ++it) {
        const uint32_t s = it % kDwStages, ph = (it / kDwStages) & 1;
        mbar_wait(&full[s], ph);
        if (bias_warp) {
          const uint32_t slab = smem_u32(smem + s * stage_stride + q * kDwSlab);
#pragma unroll 8
          for (int k = 0; k < kDwRows; ++k) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(slab + k * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)k & 7)) << 4) + ((uint32_t)lane & 3) * 4));
            bs0 += __uint_as_float(w << 16);
            bs1 += __uint_as_float(w & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      if (bias_warp) {
        const int col = J.m0 + q * 64 + lane * 2;
        if (col < J.m_valid) atomicAdd(&J.db[col], bs0);
        if (col + 1 < J.m_valid) atomicAdd(&J.db[col + 1], bs1);
      }
      // ---- dW accumulators -> global
      mbar_wait(done, 0);
      tc_fence_after();
      for (int mh = 0; mh < J.m_halves; ++mh) {
        const int m = J.m0 + mh * 128 + r;
#pragma unroll 1
        for (int c0 = 0; c0 < n_tile; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mh * n_tile + c0, v);
          tmem_ld_wait();
          if (m < J.m_valid) {
            // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic operations of the scalar form;
            // the ~16 CTAs of a job all add into the same [out, in] matrix at the end of the kernel, and that tail was 8 %
            // of the launch.  ld_w is a multiple of 4 and columns up to the next multiple of 4 past n_valid are padding.
            float* o = J.dW + (size_t)m * J.ld_w + J.n0 + c0;
#pragma unroll
            for (int e = 0; e < 32; e += 4)
              if (J.n0 + c0 + e < J.n_valid)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + e), "f"(__uint_as_float(v[e])),
                             "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                             : "memory");
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

}  // namespace tc
}  // namespace marf
