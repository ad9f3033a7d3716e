// The backward pass of the bf16 path as ONE launch (sm_100a): the dX chain and every dW / db GEMM run side by side.
//
//   k_tc_bwd : a grid of CTA pairs (clusters of 2, tcgen05 cta_group::2 everywhere).  The first `n_chain_clusters` pairs run the
//              layer-fused dX chain of tc_chain.cuh (dlogits -> dY_3 -> ... -> dY_0, activations resident in SMEM) and publish
//              every unit's output tile through a per-tile flag once its TMA stores have completed.  The other pairs are
//              "dW pairs": each owns ONE weight-gradient GEMM dW_l = dY_l^T X_l (contraction over pixel rows, both operands
//              MN-major straight from the row-major activations), keeps its 256 x 256 fp32 accumulator in the TMEM of the
//              pair for the whole launch, and consumes the dY_l tiles as the chain pairs publish them — out of L2, not HBM —
//              while X_l streams from HBM once.  db_l = column sums of dY_l falls out of a second, 16-column MMA against a
//              constant tile of ones (no SIMT pass over the stage).
//
// Before (round 1): chain launch (writes dY_l, 0.82 GB) -> dW launch (reads dY_l + X_l, 1.84 GB, HBM-bound at 0.87 of the
// copy bandwidth).  Now the dW GEMMs overlap the chain and read dY from L2.  Work split: chain pairs / dW pairs in
// proportion to their tensor work (bf16_path.cu: launch_bwd).
//
// Nothing waits on a dW pair, and a chain pair never waits on anybody, so the launch cannot deadlock as long as every CTA
// eventually becomes resident (grid <= #SMs, one CTA per SM); every wait is bounded and traps instead of hanging.
#pragma once
#include "tc_chain.cuh"

namespace marf {
namespace tc {

constexpr int kBwdMaxJobs = 12;
constexpr int kBwStages = 3;
constexpr int kBwRows = 128;                      // pixel rows per stage = one tile of the chain pairs (one hand-over flag)
constexpr int kBwSlab = kBwRows * 128;            // [128 pixel rows x 64 columns] bf16: one TMA box (16 KB; 8 KB boxes were TMA-issue-bound)
constexpr int kBwStage = 4 * kBwSlab;             // per CTA: 2 slabs of dY (its 128 output features) + 2 slabs of X (its 128 inputs)
constexpr int kBwOnesOff = kBwStages * kBwStage;  // 8 KB of bf16 1.0 (B operand of the db MMA)
constexpr int kBwBarOff = kBwOnesOff + 8192;
constexpr int kBwSmem = kBwBarOff + 256;
static_assert(kBwSmem <= kChSmem, "the dW-pair role must fit the chain role's shared memory");

struct alignas(64) BwdDwJob {
  CUtensorMap tmDY;       // box {64, 128} over dY_l [rows, out features]  (output layer: the 8-column bf16 dlogits)
  CUtensorMap tmX;        // box {64, 128} over X_l  [rows, in features]
  const uint32_t* ready;  // per-tile flags of the chain unit that produces dY_l (nullptr: complete before the launch)
  int n_cols;             // MMA N: 256, or 128 for the 64-wide inputs (the second CTA's half is TMA out-of-bounds zero fill)
  int m_valid, n_valid;   // real out / in features
  int dy_cols, x_cols;    // columns the dY / X tensors really have (256, 8 for the dlogits tile, 64 for the encoded inputs): boxes
                          // that lie wholly beyond them are not loaded (their SMEM slabs stay zero)
  int ld_w;
  int do_bias;
  float* dW;              // [out, ld_w] fp32, accumulated with red.add
  float* db;              // [out] fp32
  const unsigned char* x_base;   // X_l as raw bytes [rows, x_pitch] when this job's X tiles are loaded with cp.async by warps 2..5
  int x_pitch;                   // (256-column X only) instead of TMA: a second, independent path into SMEM (nullptr: TMA)
  unsigned char* dy_base; // dY_l as raw bytes [rows, dy_pitch] when the tiles may be DISCARDED from L2 once consumed (nullptr: keep)
  int dy_pitch;
  int pair_begin, pair_count;   // dW pairs [pair_begin, pair_begin + pair_count) take this job's 128-row stages round-robin
};
struct BwdJobs {
  ChainJobs chain;        // CH_DX jobs (ready / epoch / interleave set)
  BwdDwJob dw[kBwdMaxJobs];
  int n_dw;
  int n_chain_clusters;
  int rows;               // padded pixel rows of the chunk (multiple of 128)
  int prefetch_ahead;     // stages of X_l a dW pair prefetches into L2 ahead of its SMEM ring (0: off)
  unsigned long long* trace;   // diagnostics (MARF_BWD_TRACE): per cluster {globaltimer at role begin, at role end}; nullptr in production
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

// One dW pair.  Warp roles: warp 0 = TMA producer (one lane, in both CTAs: each CTA loads its own halves of the operands),
// warp 1 = dW MMA issuer (leader CTA) + TMEM allocation, warp 6 = db MMA issuer (leader CTA), warps 2..5 = final read-out of the
// accumulators (both CTAs).
__device__ __forceinline__ void dw_pair_role(const BwdJobs& jobs, uint8_t* smem, const int pair) {
  int jr = 0;
  for (int i = 1; i < jobs.n_dw; ++i)
    if (pair >= jobs.dw[i].pair_begin) jr = i;
  const BwdDwJob& J = jobs.dw[jr];
  const int k = pair - J.pair_begin;                 // this pair's index inside the job
  const int n_stage_tot = jobs.rows / kBwRows;
  const int n_my = (k < J.pair_count && k < n_stage_tot) ? (n_stage_tot - k + J.pair_count - 1) / J.pair_count : 0;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwBarOff);
  uint64_t* full = bars;                    // [kBwStages]  (leader: the bytes of BOTH CTAs' loads)
  uint64_t* empty = bars + kBwStages;       // [kBwStages]  (multicast commit: the MMAs that read the stage have completed)
  uint64_t* done = bars + 2 * kBwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  volatile int* progress = reinterpret_cast<volatile int*>(done + 2);       // the producer's current stage (this CTA)
  const int b_slabs = J.n_cols / 128;       // X slabs this CTA holds (2, or 1 for the 64-wide inputs)
  // A TMA box costs its 128 rows of TMA-engine time whether it fetches anything or not (measured: the operand loads bound a dW
  // pair at ~3.3 cycles per 128-byte box row), so boxes that lie wholly outside their tensor are skipped: the slabs are zeroed
  // once and never written.  real_boxes(r): how many boxes CTA r of the pair loads per stage.
  auto dy_real = [&](uint32_t r, int sl) { return (int)r * 128 + sl * 64 < J.dy_cols; };
  auto x_real = [&](uint32_t r, int b) { return ((int)r * b_slabs + b) * 64 < J.x_cols; };
  int boxes_pair = 0;
  for (uint32_t r = 0; r < 2; ++r) {
    for (int sl = 0; sl < 2; ++sl) boxes_pair += dy_real(r, sl) ? 1 : 0;
    for (int b = 0; b < b_slabs; ++b) boxes_pair += (!J.x_base && x_real(r, b)) ? 1 : 0;
  }
  const uint32_t stage_bytes_pair = (uint32_t)boxes_pair * kBwSlab;

  if (threadIdx.x == 0) {
    prefetch_tmap(&J.tmDY);
    prefetch_tmap(&J.tmX);
    // full: the producer's arrive (+ TMA bytes) and, with the cp.async path for X, one arrive per loader warp of both CTAs
    for (int s = 0; s < kBwStages; ++s) { mbar_init(&full[s], J.x_base ? 9 : 1); mbar_init(&empty[s], 2); }    // (empty: one commit per issuer)
    mbar_init(done, 2);
    *progress = 0;
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  for (int i = threadIdx.x; i < kBwStages * kBwStage / 16; i += kChThreads)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  // the constant ones tile (any layout of an all-ones slab is an all-ones operand)
  for (int i = threadIdx.x; i < 8192 / 16; i += kChThreads)
    reinterpret_cast<uint4*>(smem + kBwOnesOff)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async_smem();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: everything below consumes the previous kernels' output
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_my > 0) {
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (both CTAs; lane 0 issues)
      // Hand-over flags are read by the WHOLE warp, 32 upcoming stages at a time: one L2 round trip per batch.  (One acquire
      // load per stage in the issuing thread serialises a ~0.7 us round trip into every stage: measured 1.0 us per stage.)
      int n_ok = J.ready ? 0 : n_my;                  // stages [0, n_ok) of this pair are known to be published (warp-uniform)
      for (int i = 0; i < n_my; ++i) {
        if (i >= n_ok) {
          uint32_t spins = 0;
          for (;;) {
            const int ii = i + lane;
            bool ok = true;
            if (ii < n_my) ok = (int32_t)(ld_acquire_gpu(J.ready + (k + ii * J.pair_count)) - jobs.chain.epoch) >= 0;
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            const int cnt = m == 0xffffffffu ? 32 : __ffs((int)~m) - 1;      // leading published stages of the batch
            if (cnt > 0) { n_ok = i + cnt; break; }
            __nanosleep(256);
            if (++spins > (1u << 22)) __trap();
          }
          __syncwarp();                                // the lanes' acquires are ordered before lane 0's TMA issue ...
          if (lane == 0) fence_proxy_async_all();      // ... and before the async-proxy (TMA) reads.  Once per batch: this fence
        }                                              // costs ~0.35 us (measured: 0.79 -> 0.43 us per stage when issued per stage)
        if (lane == 0) {
          const uint32_t s = i % kBwStages, ph = (i / kBwStages) & 1;
          const int row = (k + i * J.pair_count) * kBwRows;
          *progress = i;                               // (paces the L2 prefetcher, warp 7)
          mbar_wait(&empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&full[s], stage_bytes_pair);
          uint8_t* st = smem + s * kBwStage;
          // dY: this CTA's 128 output features; X: its half of the N columns (boxes partly beyond the tensor's width: zero fill)
          for (int sl = 0; sl < 2; ++sl)
            if (dy_real(rank, sl)) tma_load_2d_2sm(st + sl * kBwSlab, &J.tmDY, (int)rank * 128 + sl * 64, row, &full[s], kEvictFirst);
          if (!J.x_base)
            for (int b = 0; b < b_slabs; ++b)
              if (x_real(rank, b)) tma_load_2d_2sm(st + (2 + b) * kBwSlab, &J.tmX, ((int)rank * b_slabs + b) * 64, row, &full[s], kEvictFirst);
        }
        __syncwarp();
      }
    } else if (warp == 1 || warp == 6) {
      // ---------------------------------------------------------------- MMA issuers (leader CTA): warp 1 the dW MMAs, warp 6
      // the db MMAs.  One thread issues an MMA every ~90 cycles whatever its size, so the four 16-column db MMAs of a stage
      // would cost the dW stream as much issue time as its own four 256-column MMAs (measured: 810 cycles per stage with one
      // issuer against 512 cycles of tensor work).  The accumulators are independent; every MMA-side barrier takes both commits.
      if (lane == 0 && rank == 0) {
        const bool bias_issuer = warp == 6;
        const uint32_t idesc = bias_issuer ? idesc_bf16(256, 16, 1, 1) : idesc_bf16(256, J.n_cols, 1, 1);   // both operands MN-major
        const uint32_t d_tmem = tmem_base + (bias_issuer ? 256u : 0u);
        const uint32_t s_ones = smem_u32(smem + kBwOnesOff);
        for (int i = 0; i < n_my; ++i) {
          const uint32_t s = i % kBwStages, ph = (i / kBwStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * kBwStage);
          const uint32_t b0 = bias_issuer ? s_ones : st + 2 * kBwSlab;          // (arithmetic select: no branch around the MMAs)
          const uint32_t bstep = bias_issuer ? 0u : 2048u;                       // (the ones tile is the same for every K step)
#pragma unroll
          for (int ks = 0; ks < kBwRows / 16; ++ks) {
            const uint64_t da = smem_desc_sw128(st + ks * 2048, kBwSlab, 1024);
            const uint64_t db = smem_desc_sw128(b0 + ks * bstep, kBwSlab, 1024);
            umma_bf16_2sm(d_tmem, da, db, idesc, (i | ks) != 0);
          }
          umma_commit_2sm(&empty[s]);
        }
        umma_commit_2sm(done);
      }
    } else if (warp == 7) {
      // ---------------------------------------------------------------- L2 prefetcher (both CTAs): X_l has been in HBM since the
      // forward pass; this thread pulls the pair's X tiles into L2 `prefetch_ahead` stages ahead of the producer, so that the SMEM
      // ring (192 KB per CTA, all there is) sees L2 latency instead of loaded HBM latency.  Its own thread: a prefetch issued by
      // the producer thread delays that thread's loads (measured).
      const int pf = jobs.prefetch_ahead;
      if (lane == 0 && pf > 0) {
        int j = 0;
        while (j < n_my) {
          const int lim = min(n_my, *progress + pf);
          for (; j < lim; ++j) {
            const int row = (k + j * J.pair_count) * kBwRows;
            for (int b = 0; b < b_slabs; ++b) tma_prefetch_2d(&J.tmX, ((int)rank * b_slabs + b) * 64, row);
          }
          if (j < n_my) __nanosleep(500);
        }
      }
    } else if (warp < 6) {
      // ---------------------------------------------------------------- warps 2..5 (both CTAs): X loader, L2 discard, read-out
      const int q = warp & 3;
      const int r = q * 32 + lane;
      const int m = (int)rank * 128 + r;
      // (1) X tiles by cp.async (LDGSTS): the TMA engine delivers ~40 B/clk per SM however the boxes are shaped (3.3 cycles per
      //     128-byte box row), which bounds a dW pair at 0.85 us per 128-row stage against 0.54 us of MMAs; 128 threads copying
      //     16-byte chunks straight into the SW128 slabs are a second path into SMEM that runs beside it.  Thread = one 16-byte
      //     column chunk of the CTA's 256-byte half row, 16 rows per stage (rows tid/16 + 8 j).
      // (2) Nobody reads a dY tile again once its MMAs have completed: drop this CTA's half of it (128 rows x 256 B) from L2
      //     without a write-back (discard.global.L2), so that the hand-over costs no DRAM write either.
      const int tid = (warp - 2) * 32 + lane;
      const int ch = tid & 15, row0 = tid >> 4;                       // chunk in the half row (16 x 16 B), first row
      const uint32_t full_leader0 = mapa_u32(smem_u32(&full[0]), 0);
      auto discard_stage = [&](int i) {
        unsigned char* p = J.dy_base + (size_t)((k + i * J.pair_count) * kBwRows + r) * J.dy_pitch + rank * 256;
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + 128) : "memory");
      };
      auto arrive_full = [&](int i) {                                  // this warp's chunks of stage i are in SMEM
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader0 + 8 * (uint32_t)(i % kBwStages));
      };
      for (int i = 0; i < n_my; ++i) {
        const uint32_t s = i % kBwStages, ph = (i / kBwStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);                                  // the MMAs of stage i - 3 have completed: slot free,
        if (J.dy_base && i >= kBwStages) discard_stage(i - kBwStages); // and its dY tile is dead
        if (J.x_base) {
          const unsigned char* g = J.x_base + (size_t)((k + i * J.pair_count) * kBwRows + row0) * J.x_pitch + rank * 256 + ch * 16;
          const uint32_t dst = smem_u32(smem + s * kBwStage + (2 + (ch >> 3)) * kBwSlab);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int row = row0 + 8 * j;
            const uint32_t a = dst + (uint32_t)(row >> 3) * 1024 + (uint32_t)(row & 7) * 128 + (((uint32_t)(ch & 7) ^ (uint32_t)(row & 7)) << 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g + (size_t)(8 * j) * J.x_pitch) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (i >= 1) {                                                // two stages of copies in flight per thread
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            arrive_full(i - 1);
          }
        }
      }
      if (J.x_base) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        arrive_full(n_my - 1);
      }
      if (J.dy_base)
        for (int i = max(0, n_my - kBwStages); i < n_my; ++i) {
          mbar_wait(&empty[i % kBwStages], (i / kBwStages) & 1);
          discard_stage(i);
        }
      mbar_wait(done, 0);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
      if (J.do_bias) {
        uint32_t v;
        tmem_ld1(t_lane + 256, v);
        tmem_ld_wait();
        if (m < J.m_valid) atomicAdd(&J.db[m], __uint_as_float(v));
      }
#pragma unroll 1
      for (int c0 = 0; c0 < J.n_cols; c0 += 32) {
        if (c0 >= J.n_valid) break;                  // (warp-uniform)
        uint32_t v[32];
        tmem_ld32(t_lane + c0, v);
        tmem_ld_wait();
        if (m < J.m_valid) {
          // 16-byte vector reductions; ld_w is a multiple of 4 and columns up to the next multiple of 4 past n_valid are padding
          float* o = J.dW + (size_t)m * J.ld_w + c0;
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            if (c0 + e < J.n_valid)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + e), "f"(__uint_as_float(v[e])),
                           "f"(__uint_as_float(v[e + 1])), "f"(__uint_as_float(v[e + 2])), "f"(__uint_as_float(v[e + 3]))
                           : "memory");
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int SCHED>
__global__ void __launch_bounds__(kChThreads, 1) k_tc_bwd(const __grid_constant__ BwdJobs jobs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int cluster = (int)blockIdx.x / 2;
  if (jobs.trace && threadIdx.x == 0 && (blockIdx.x & 1) == 0) jobs.trace[2 * cluster] = globaltimer_ns();
  if (cluster < jobs.n_chain_clusters) {
    if (SCHED == 1) chain_role_staggered<CH_DX>(jobs.chain, smem, cluster, jobs.n_chain_clusters);
    else chain_role<CH_DX, 2>(jobs.chain, smem, cluster, jobs.n_chain_clusters);
  } else dw_pair_role(jobs, smem, cluster - jobs.n_chain_clusters);
  if (jobs.trace && threadIdx.x == 0 && (blockIdx.x & 1) == 0) jobs.trace[2 * cluster + 1] = globaltimer_ns();
}

}  // namespace tc
}  // namespace marf
