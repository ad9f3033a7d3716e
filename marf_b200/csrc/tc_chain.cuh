// Layer-fused tcgen05 chain kernel of the bf16 path (sm_100a).
//
//   k_tc_chain<CH_FWD> : X_{l+1} = relu(X_l W_l^T + b_l) for the four 256-wide hidden layers of an MLP, then the
//                        3-/1-wide output layer, for a PAIR of 128-row tiles per pass.  The activations of the pair stay
//                        in SMEM between the layers (written in place by the epilogue, in the SWIZZLE_128B K-major
//                        layout the next layer's MMA reads); every layer's output is still TMA-stored to HBM once
//                        because the dW pass needs it, but nothing is read back: HBM traffic per pixel-sample is
//                        128 B in + 4 x 512 B out instead of 4 x (512 B in + 512 B out).
//   k_tc_chain<CH_DX>  : dY_{l-1} = (dY_l W_l) * relu_mask(X_l) from the output-layer gradient down to dY_0, same scheme.
//
// The weights do not fit in SMEM next to the activations, so they stream from L2 through a 4-stage TMA ring as
// [128 output features x 64 K] chunks; one chunk feeds the MMAs of both tiles (so 64 KB of weights per tile-layer).
// A layer is issued as two output halves h (128 columns each, accumulators acc[tile][h] = the 512 TMEM columns):
// while the epilogue drains half h, the MMAs of the other half / of the next layer run.
//
// Warp roles (352 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..9 = epilogue
// (group g = (warp-2)/4 owns tile g of the pair, warp%4 = TMEM lane quarter, thread = tile row), warp 10 = second MMA
// issuer of the CTA-pair variant (idle otherwise).
#pragma once
#include "tc_kernels.cuh"

namespace marf {
namespace tc {

enum { CH_FWD = 0, CH_DX = 1 };

constexpr int kChThreads = 352;
constexpr int kChUnits = 4;                   // hidden units of a chain: 64 -> 256 -> 256 -> 256 -> 256
constexpr int kChWStages = 3;
constexpr int kChWStage = 128 * 128;          // [128 features x 64 K] bf16
constexpr int kChSlab = kChunkBytes;          // [128 rows x 64 cols] bf16 = 16 KB
constexpr int kChActOff = 0;                  // act[2 tiles][4 slabs]
constexpr int kChWOff = 2 * 4 * kChSlab;      // W ring
constexpr int kChInOff = kChWOff + kChWStages * kChWStage;   // in[2 tiles]: the 64-wide input of unit 0
constexpr int kChBiasOff = kChInOff + 2 * kChSlab;           // CH_FWD: bias[2 chains][4 units][256] fp32, resident
constexpr int kChBarOff = kChBiasOff + 2 * kChUnits * 256 * 4;
constexpr int kChSmem = kChBarOff + 256;

struct alignas(64) ChainUnit {
  CUtensorMap tmW;        // box {64, 128} over the unit's weights [256 features, K]
  CUtensorMap tmOut;      // box {64, 128} over the unit's output [rows, 256]
  const float* bias;      // CH_FWD: [256]
  uint32_t* bits;         // CH_FWD: ReLU mask of the output (written); CH_DX: ReLU mask applied to the output (read)
};
struct alignas(64) ChainJob {
  CUtensorMap tmIn;       // box {64, 128} over the 64-wide input of unit 0 [rows, 64]
  CUtensorMap tmWout;     // CH_FWD: box {64, 16} over the output layer's [16, 256] hi/lo weight rows
  ChainUnit u[kChUnits];
  const float* bias_out;  // CH_FWD: [k_out]
  float* logits;          // CH_FWD: [rows, 4] fp32
  int k_out;              // CH_FWD: 3 or 1
  int bits_ld;            // words per row of the mask arrays (= 8)
};
struct ChainJobs {
  ChainJob c[2];
  int n;                  // chains in this launch
  int n_tiles;            // 128-row tiles per chain
  int dbg;                // timing experiments only (results invalid): 1 no TMA stores, 2 weights loaded once per CTA, 4 no epilogue math
  long long* trace;       // diagnostics: clock64() stamps of CTA 0, [item < 2][unit][half][16]; nullptr in production
  // CH_DX inside k_tc_bwd (tc_bwd.cuh): the dW pairs of the same launch consume every unit's output tile by tile.
  uint32_t* ready;        // [chain][unit][n_tiles]: set to `epoch` (release, gpu scope) once the tile's TMA stores have completed
  uint32_t epoch;         // launch counter of the handle (flags are never reset)
  int store_last;         // dY tiles are stored with L2 evict_last priority (they are consumed out of L2 by the same launch)
  int interleave;         // work item -> (chain = item % n, tile group = item / n) instead of chain-major order
};

__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
// (no memory clobber, not volatile: the bias table is written once before the role split, the compiler may hoist these)
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// CL = 1: one CTA per tile pair.  CL = 2: a CTA pair (cluster of 2, cta_group::2): the two CTAs hold two tiles each, every
// MMA covers 256 rows (tile t of both CTAs), each CTA loads and holds only HALF of every weight chunk (its 64 of the 128
// output features), the leader CTA (cluster rank 0) issues all MMAs and its commits are multicast to both CTAs; the
// epilogue warps of both CTAs arrive on the leader's epi_done barriers.  Per tile-layer this takes a quarter of the SMEM
// bandwidth off the MMA operand reads and halves the weight stream.
template <int MODE, int CL>
__device__ __forceinline__ void chain_role(const ChainJobs& jobs, uint8_t* smem, const int cid, const int ncl) {
  constexpr bool kOut = MODE == CH_FWD;          // the forward chain ends with the thin output layer
  const uint32_t s_act = smem_u32(smem + kChActOff);
  const uint32_t s_w = smem_u32(smem + kChWOff);
  const uint32_t s_in = smem_u32(smem + kChInOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChBarOff);
  // weight ring: the 48 KB hold 3 whole chunks (CL = 1) or 6 half chunks (CL = 2: each CTA keeps only its 64 features)
  constexpr int kStages = kChWStages * CL;
  constexpr uint32_t kStageBytes = kChWStage / CL;
  uint64_t* w_full = bars;                       // [<= 6]
  uint64_t* w_empty = bars + 6;                  // [<= 6]
  uint64_t* in_full = bars + 12;
  uint64_t* in_empty = bars + 13;
  uint64_t* acc_full = bars + 14;                // [2]  (per output half)
  uint64_t* epi_done = bars + 16;                // [2]  (per output half): accumulators drained, output slabs written
  uint64_t* slab01_free = bars + 18;             // the MMAs that read slabs 0,1 of the current layer input are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL == 2 ? cluster_ctarank() : 0u;
  // work is distributed over clusters: cluster `cid` of `ncl` takes the items cid, cid + ncl, ...
  constexpr int kGroup = 2 * CL;                                        // tiles per work item
  const int n_pairs = (jobs.n_tiles + kGroup - 1) / kGroup;             // work items per chain
  const int n_items = n_pairs * jobs.n;
  const bool ilv = jobs.interleave != 0;
  auto chain_of = [&](int item) -> int { return ilv ? item % jobs.n : item / n_pairs; };
  auto group_of = [&](int item) -> int { return ilv ? item / jobs.n : item % n_pairs; };

  if (threadIdx.x == 0) {
    for (int c = 0; c < jobs.n; ++c) {
      prefetch_tmap(&jobs.c[c].tmIn);
      for (int u = 0; u < kChUnits; ++u) { prefetch_tmap(&jobs.c[c].u[u].tmW); prefetch_tmap(&jobs.c[c].u[u].tmOut); }
      if (kOut) prefetch_tmap(&jobs.c[c].tmWout);
    }
    // (MMA-side barriers take one commit from each issuer thread: CL of them)
    for (int s = 0; s < kStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], CL); }
    mbar_init(in_full, 1);
    mbar_init(in_empty, CL);
    for (int h = 0; h < 2; ++h) { mbar_init(&acc_full[h], CL); mbar_init(&epi_done[h], 8 * CL); }
    mbar_init(slab01_free, CL);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL == 2) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: everything below consumes the previous kernels' output
  if (MODE == CH_FWD) {
    float* sb = reinterpret_cast<float*>(smem + kChBiasOff);
    for (int i = threadIdx.x; i < jobs.n * kChUnits * 256; i += kChThreads)
      sb[i] = jobs.c[i >> 10].u[(i >> 8) & 3].bias[i & 255];
  }
  tc_fence_before();
  if (CL == 2) cluster_sync_all();        // the peer's barriers must be initialised before anything is signalled on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t wit = 0, n_in = 0;
      auto load_w = [&](const CUtensorMap* tm, int k, int h) {
        const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
        mbar_wait(&w_empty[s], ph ^ 1);
        if (CL == 1 && (jobs.dbg & 2) && wit >= kStages) { mbar_arrive(&w_full[s]); ++wit; return; }
        if (rank == 0) mbar_expect_tx(&w_full[s], kChWStage);           // (both CTAs' halves land on the leader's barrier)
        if (CL == 2) tma_load_2d_2sm(smem + kChWOff + s * kStageBytes, tm, k * kChunkK, h * 128 + (int)rank * 64, &w_full[s], kEvictLast);
        else tma_load_2d_hint(smem + kChWOff + s * kStageBytes, tm, k * kChunkK, h * 128, &w_full[s], kEvictLast);
        ++wit;
      };
      auto load_in = [&](int item) {
        const ChainJob& J = jobs.c[chain_of(item)];
        const int g0 = kGroup * group_of(item);                         // first tile of the work item
        const int tile0 = g0 + 2 * (int)rank;                           // this CTA's two tiles
        const int n_valid = min(kGroup, jobs.n_tiles - g0);
        mbar_wait(in_empty, (n_in & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(in_full, (uint32_t)n_valid * kChSlab);
        for (int t = 0; t < 2; ++t)
          if (tile0 + t < jobs.n_tiles) {
            if (CL == 2) tma_load_2d_2sm(smem + kChInOff + t * kChSlab, &J.tmIn, 0, (tile0 + t) * kTileM, in_full, kEvictFirst);
            else tma_load_2d_hint(smem + kChInOff + t * kChSlab, &J.tmIn, 0, (tile0 + t) * kTileM, in_full, kEvictFirst);
          }
        ++n_in;
      };
      if (cid < n_items) load_in(cid);
      for (int item = cid; item < n_items; item += ncl) {
        const ChainJob& J = jobs.c[chain_of(item)];
        load_w(&J.u[0].tmW, 0, 0);
        load_w(&J.u[0].tmW, 0, 1);
        for (int u = 1; u < kChUnits; ++u) {
          for (int h = 0; h < 2; ++h)
            for (int k = 0; k < 4; ++k) load_w(&J.u[u].tmW, k, h);
          // the next pair's input as soon as unit 0 of this pair has consumed the buffer (in_empty): long before it is needed
          if (u == 1 && item + ncl < n_items) load_in(item + ncl);
        }
        if (kOut) {
          const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          if (CL == 1 && (jobs.dbg & 2)) { mbar_arrive(&w_full[s]); ++wit; continue; }
          if (rank == 0) mbar_expect_tx(&w_full[s], 4 * 2048);
          for (int k = 0; k < 4; ++k) {          // (CL = 2: this CTA's 8 of the 16 hi/lo rows, 1 KB per K chunk)
            if (CL == 2) tma_load_2d_2sm(smem + kChWOff + s * kStageBytes + k * 1024, &J.tmWout, k * kChunkK, (int)rank * 8, &w_full[s], kEvictLast);
            else tma_load_2d_hint(smem + kChWOff + s * kStageBytes + k * 2048, &J.tmWout, k * kChunkK, 0, &w_full[s], kEvictLast);
          }
          ++wit;
        }
      }
    }
  } else if (warp == 1 || (CL == 2 && warp == 10)) {
    // ------------------------------------------------------------------ MMA issuer(s).  CL = 1: warp 1 issues for both tiles.
    // CL = 2: in the leader CTA only, warp 1 issues the MMAs of tile 0 (of both CTAs), warp 10 those of tile 1: one thread
    // cannot issue the 64 MMAs of a unit faster than ~90 cycles each; the two tiles' accumulators are independent, so the
    // two streams need no mutual order, and every MMA-side barrier takes one commit from each issuer.
    if (lane == 0 && rank == 0) {
      const uint32_t mi = warp == 1 ? 0u : 1u;
      constexpr uint32_t idesc = idesc_bf16(kTileM * CL, 128, 0, 0);
      constexpr uint32_t idesc_out = idesc_bf16(kTileM * CL, 16, 0, 0);
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if (CL == 2) umma_bf16_2sm(d, da, db, id, acc); else umma_bf16(d, da, db, id, acc);
      };
      auto commit = [&](uint64_t* bar) { if (CL == 2) umma_commit_2sm(bar); else umma_commit(bar); };
      auto wait_x = [&](uint64_t* bar, uint32_t parity) {      // barriers the peer CTA signals too (plain acquire, as CUTLASS)
        if (CL == 2 && (jobs.dbg & 8)) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
      };
      uint32_t wit = 0, n_in = 0, c_epi0 = 0, c_epi1 = 0;
      long long w_wait = 0;
      auto wait_epi = [&](int h) {
        wait_x(&epi_done[h], (h ? c_epi1 : c_epi0) & 1);
        if (h) ++c_epi1; else ++c_epi0;
        tc_fence_after();
      };
      // one weight chunk against the same K slab of both tiles
      auto mma_chunk = [&](uint32_t a0, uint32_t a1, int h, bool first) {
        const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
        const long long tw0 = jobs.trace ? clock64() : 0;
        wait_x(&w_full[s], ph);
        if (jobs.trace) w_wait += clock64() - tw0;
        tc_fence_after();
        const uint64_t db = smem_desc_sw128(s_w + s * kStageBytes, 16, 1024);
        // Every MMA unconditionally, also for tile 1 of a work item whose second tile does not exist (its accumulators are
        // then never read): a branch around tcgen05.mma is if-converted by ptxas 12.9 into predicated UTCHMMAs whose
        // descriptor R2URs hang on an unrelated predicate (observed: stale A/B descriptors, wrong results).
        if (CL == 1) {
          const uint64_t d0 = smem_desc_sw128(a0, 16, 1024);
          const uint64_t d1 = smem_desc_sw128(a1, 16, 1024);
#pragma unroll
          for (int j = 0; j < 4; ++j) mma(tmem_base + h * 128, d0 + 2 * j, db + 2 * j, idesc, !(first && j == 0));
#pragma unroll
          for (int j = 0; j < 4; ++j) mma(tmem_base + 256 + h * 128, d1 + 2 * j, db + 2 * j, idesc, !(first && j == 0));
        } else {
          const uint64_t da = smem_desc_sw128(a0 + mi * (a1 - a0), 16, 1024);      // this issuer's tile (arithmetic, no branch)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma(tmem_base + mi * 256 + h * 128, da + 2 * j, db + 2 * j, idesc, !(first && j == 0));
        }
        commit(&w_empty[s]);
        ++wit;
      };
      bool first_item = true;
      int it_no = 0;
      auto stamp = [&](int u, int h, int e) {
        if (jobs.trace && cid == 0 && rank == 0 && it_no < 2 && mi == 0) {
          jobs.trace[((it_no * 4 + u) * 2 + h) * 16 + e] = clock64();
          if (e == 1) jobs.trace[((it_no * 4 + u) * 2 + h) * 16 + 3] = jobs.trace[0] + w_wait;   // cumulative wait on weight chunks
        }
      };
      for (int item = cid; item < n_items; item += ncl, first_item = false, ++it_no) {
        // ---- unit 0: K = 64 from the input buffer
        wait_x(in_full, n_in & 1);
        ++n_in;
        tc_fence_after();
        if (!first_item && !kOut) wait_epi(0);          // acc[.][0] drained (last unit of the previous pair)
        stamp(0, 0, 0);
        mma_chunk(s_in, s_in + kChSlab, 0, true);
        commit(&acc_full[0]);
        stamp(0, 0, 1);
        if (!first_item) wait_epi(1);                   // acc[.][1] drained (output unit / last unit of the previous pair)
        stamp(0, 1, 0);
        mma_chunk(s_in, s_in + kChSlab, 1, true);
        commit(&acc_full[1]);
        stamp(0, 1, 1);
        commit(in_empty);
        // ---- units 1..3: K = 256 from the activation slabs the previous unit's epilogue wrote in place
        for (int u = 1; u < kChUnits; ++u) {
          wait_epi(0);                                  // slabs 0,1 written, acc[.][0] drained
          stamp(u, 0, 0);
          mma_chunk(s_act + 0 * kChSlab, s_act + 4 * kChSlab, 0, true);
          mma_chunk(s_act + 1 * kChSlab, s_act + 5 * kChSlab, 0, false);
          wait_epi(1);                                  // slabs 2,3 written, acc[.][1] drained
          stamp(u, 0, 2);
          mma_chunk(s_act + 2 * kChSlab, s_act + 6 * kChSlab, 0, false);
          mma_chunk(s_act + 3 * kChSlab, s_act + 7 * kChSlab, 0, false);
          commit(&acc_full[0]);
          stamp(u, 0, 1);
          stamp(u, 1, 0);
          mma_chunk(s_act + 0 * kChSlab, s_act + 4 * kChSlab, 1, true);
          mma_chunk(s_act + 1 * kChSlab, s_act + 5 * kChSlab, 1, false);
          commit(slab01_free);                     // the half-0 epilogue may now overwrite slabs 0,1
          mma_chunk(s_act + 2 * kChSlab, s_act + 6 * kChSlab, 1, false);
          mma_chunk(s_act + 3 * kChSlab, s_act + 7 * kChSlab, 1, false);
          commit(&acc_full[1]);
          stamp(u, 1, 1);
        }
        if (kOut) {
          // ---- output layer: N = 16 (rows 0..7 hi, 8..15 lo halves of the <= 4 real output rows), accumulators in acc[.][1]
          wait_epi(0);
          wait_epi(1);
          const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
          wait_x(&w_full[s], ph);
          tc_fence_after();
#pragma unroll
          for (int tt = 0; tt < (CL == 1 ? 2 : 1); ++tt) {
            const uint32_t t = CL == 1 ? (uint32_t)tt : mi;          // CL = 2: this issuer's tile
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = smem_desc_sw128(s_act + (t * 4 + k) * kChSlab, 16, 1024);
              const uint64_t db = smem_desc_sw128(s_w + s * kStageBytes + k * (2048 / CL), 16, 1024);
#pragma unroll
              for (int j = 0; j < 4; ++j) mma(tmem_base + t * 256 + 128, da + 2 * j, db + 2 * j, idesc_out, (k | j) != 0);
            }
          }
          commit(&w_empty[s]);
          ++wit;
          commit(&acc_full[1]);
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue: group = tile of the pair, thread = row
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const bool gleader = threadIdx.x == 64 + 128 * grp;
    const uint32_t sw_row = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 256;
    uint32_t c_acc0 = 0, c_acc1 = 0, c_free = 0;
    int it_no = 0;
    uint32_t* pend_flag = nullptr;       // (group leader) flag of the newest stored tile that is not published yet
    const uint32_t epi_bar0 = CL == 2 ? mapa_u32(smem_u32(&epi_done[0]), 0) : 0u;     // the leader CTA's epi_done[0]
    for (int item = cid; item < n_items; item += ncl, ++it_no) {
      const bool tr = jobs.trace && cid == 0 && rank == 0 && it_no < 2 && (threadIdx.x == 64 || threadIdx.x == 192);
      auto stamp = [&](int u, int h, int e) {
        if (tr) jobs.trace[((it_no * 4 + u) * 2 + h) * 16 + 4 + grp * 6 + e] = clock64();
      };
      const int chain = chain_of(item);
      const ChainJob& J = jobs.c[chain];
      const int tile = kGroup * group_of(item) + 2 * (int)rank + grp;
      const bool valid = tile < jobs.n_tiles;
      const size_t grow = (size_t)tile * kTileM + r;
      for (int u = 0; u < kChUnits; ++u) {
        const ChainUnit& U = J.u[u];
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint4 mb = make_uint4(0u, 0u, 0u, 0u);
          if (MODE == CH_DX && valid) mb = *reinterpret_cast<const uint4*>(U.bits + grow * J.bits_ld + h * 4);
          mbar_wait(&acc_full[h], (h ? c_acc1 : c_acc0) & 1);
          if (h) ++c_acc1; else ++c_acc0;
          tc_fence_after();
          stamp(u, h, 0);
          if (valid) {
            // the TMA stores that read slabs 2h, 2h+1 (this group's second-newest commit group) must have finished reading
            if (gleader) bulk_wait_read<1>();
            named_bar_sync(1 + grp, 128);
          }
          if (u >= 1 && h == 0) {
            mbar_wait(slab01_free, c_free & 1);
            ++c_free;
          }
          stamp(u, h, 1);
          if (valid && !(jobs.dbg & 4)) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = 2 * h + jj;
              uint32_t v[64];
              {
                uint32_t (&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                uint32_t (&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                tmem_ld32(t_lane + h * 128 + jj * 64, v0);
                tmem_ld32(t_lane + h * 128 + jj * 64 + 32, v1);
              }
              const uint32_t ob = s_act + (uint32_t)(grp * 4 + j) * kChSlab + sw_row;
              uint32_t obits[2] = {0u, 0u};
              const uint32_t mw[2] = {jj == 0 ? mb.x : mb.z, jj == 0 ? mb.y : mb.w};
              const uint32_t bp = smem_u32(smem + kChBiasOff) + (uint32_t)((chain * kChUnits + u) * 256 + j * 64) * 4;
              tmem_ld_wait();
              stamp(u, h, 2 + jj);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                uint32_t w[4];
                float bv[8];
                if (MODE == CH_FWD) {
                  const float4 b0 = lds128f(bp + i * 32), b1 = lds128f(bp + i * 32 + 16);
                  bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
                }
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                  const int c = i * 8 + pr * 2;             // column inside the slab
                  const int t = (c & 31) >> 1;              // pair index inside the 32-column group
                  float lo = __uint_as_float(v[c]), hi = __uint_as_float(v[c + 1]);
                  if (MODE == CH_FWD) {
                    lo += bv[pr * 2];
                    hi += bv[pr * 2 + 1];
                    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                    uint32_t gt;
                    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(gt) : "r"(w[pr]), "r"(0u));
                    obits[c >> 5] |= gt & (0x00010001u << t);
                  } else {
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                    const uint32_t sel = (mw[c >> 5] >> t) & 0x00010001u;
                    w[pr] &= sel * 0xFFFFu;
                  }
                }
                sts128(ob + (((uint32_t)i ^ (uint32_t)(r & 7)) << 4), w[0], w[1], w[2], w[3]);
              }
              if (MODE == CH_FWD)
                *reinterpret_cast<uint2*>(U.bits + grow * J.bits_ld + j * 2) = make_uint2(obits[0], obits[1]);
            }
            fence_proxy_async_smem();
          }
          stamp(u, h, 4);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {                                 // the MMA warp may read the slabs / overwrite the accumulators
            if (CL == 2) mbar_arrive_cluster(epi_bar0 + 8 * h); else mbar_arrive(&epi_done[h]);
          }
          if (valid) {
            named_bar_sync(3 + grp, 128);                  // all 128 rows of both slabs are in SMEM
            if (gleader && !(jobs.dbg & 1)) {
              // forward activations are next read by the dW pass, a whole backward pass later: evict first.  The dY of the dX
              // chain are read by the dX0 GEMM and the dW launch right after this kernel: normal priority, so that what the L2
              // still holds of them at the end of the kernel is served from there.
              const uint64_t pol = MODE == CH_FWD ? kEvictFirst : (jobs.store_last ? kEvictLast : kEvictNormal);
              tma_store_2d_hint(&U.tmOut, (2 * h) * 64, tile * kTileM, smem + kChActOff + (grp * 4 + 2 * h) * kChSlab, pol);
              tma_store_2d_hint(&U.tmOut, (2 * h + 1) * 64, tile * kTileM, smem + kChActOff + (grp * 4 + 2 * h + 1) * kChSlab, pol);
              bulk_commit();
              if (MODE == CH_DX && jobs.ready && h == 1) {
                // lagged hand-over to the dW pairs: everything older than this unit's two store groups has completed, i.e.
                // the tile this group stored one unit ago (a TMA store needs ~3k cycles; waiting for the newest would stall)
                if (pend_flag) {
                  bulk_wait<2>();
                  fence_proxy_async_all();
                  st_release_gpu(pend_flag, jobs.epoch);
                }
                pend_flag = jobs.ready + ((size_t)(chain * kChUnits + u) * jobs.n_tiles + tile);
              }
            }
          }
          stamp(u, h, 5);
        }
      }
      if (kOut) {
        mbar_wait(&acc_full[1], c_acc1 & 1);
        ++c_acc1;
        tc_fence_after();
        if (valid) {
          uint32_t v[16];
          tmem_ld16(t_lane + 128, v);
          tmem_ld_wait();
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = e < J.k_out ? __uint_as_float(v[e]) + __uint_as_float(v[8 + e]) + __ldg(J.bias_out + e) : 0.f;
          *reinterpret_cast<float4*>(J.logits + grow * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CL == 2) mbar_arrive_cluster(epi_bar0 + 8); else mbar_arrive(&epi_done[1]);
        }
      }
    }
    if (gleader) {
      bulk_wait<0>();
      if (MODE == CH_DX && pend_flag) {
        fence_proxy_async_all();
        st_release_gpu(pend_flag, jobs.epoch);
      }
    }
  }
  tc_fence_before();
  if (CL == 2) cluster_sync_all();        // the peer may still signal this CTA's barriers / read its SMEM until here
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CL == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================================
// The same chain on CTA pairs with a TILE-STAGGERED, full-width schedule (SCHED = 1):
//   every layer of a tile is ONE accumulator of 256 columns (acc[tile] = 256 of the CTA's 512 TMEM columns) filled by N = 256 MMAs,
//   and the two tiles of a CTA alternate: while the epilogue group of tile 0 drains and rewrites tile 0's slabs, the MMAs of tile 1
//   run, and vice versa.  Against the two-halves schedule above (both tiles in lockstep, N = 128 halves) this
//     * reads the A operand once per layer instead of once per half (SMEM: 352 -> 320 KB per tile-layer),
//     * has no layer-boundary bubble (the other tile's MMAs fill it) and needs no slab hand-over barrier (all MMAs that read a
//       tile's slabs have completed when its accumulator is complete),
//     * issues a quarter of the MMA instructions (one issuer thread suffices),
//   and pays with streaming every layer's weights twice per work item (once per tile; L2 -> SMEM 128 KB per layer per CTA).
// Barriers: w_full/w_empty [3 stages of 16 KB: this CTA's 128 of the 256 weight rows x 64 K], in_full/in_empty, acc_full[tile],
// epi_done[tile] (the tile's slabs are rewritten and its accumulator is drained: 4 warps of each CTA).
template <int MODE>
__device__ __forceinline__ void chain_role_staggered(const ChainJobs& jobs, uint8_t* smem, const int cid, const int ncl) {
  constexpr bool kOut = MODE == CH_FWD;
  constexpr int kStages = 3;
  constexpr uint32_t kStageBytes = 128 * 128;       // [128 weight rows x 64 K] bf16 per CTA
  const uint32_t s_act = smem_u32(smem + kChActOff);
  const uint32_t s_w = smem_u32(smem + kChWOff);
  const uint32_t s_in = smem_u32(smem + kChInOff);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kChBarOff);
  uint64_t* w_full = bars;                       // [3]
  uint64_t* w_empty = bars + 3;                  // [3]
  uint64_t* in_full = bars + 6;
  uint64_t* in_empty = bars + 7;
  uint64_t* acc_full = bars + 8;                 // [2]  (per tile)
  uint64_t* epi_done = bars + 10;                // [2]  (per tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr int kGroup = 4;                                             // tiles per work item (two per CTA)
  const int n_pairs = (jobs.n_tiles + kGroup - 1) / kGroup;
  const int n_items = n_pairs * jobs.n;
  const bool ilv = jobs.interleave != 0;
  auto chain_of = [&](int item) -> int { return ilv ? item % jobs.n : item / n_pairs; };
  auto group_of = [&](int item) -> int { return ilv ? item / jobs.n : item % n_pairs; };

  if (threadIdx.x == 0) {
    for (int c = 0; c < jobs.n; ++c) {
      prefetch_tmap(&jobs.c[c].tmIn);
      for (int u = 0; u < kChUnits; ++u) { prefetch_tmap(&jobs.c[c].u[u].tmW); prefetch_tmap(&jobs.c[c].u[u].tmOut); }
      if (kOut) prefetch_tmap(&jobs.c[c].tmWout);
    }
    for (int s = 0; s < kStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(in_full, 1);
    mbar_init(in_empty, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&acc_full[t], 1); mbar_init(&epi_done[t], 8); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc_2sm(tmem_slot, 512); tmem_relinquish_2sm(); }
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: everything below consumes the previous kernels' output
  if (MODE == CH_FWD) {
    float* sb = reinterpret_cast<float*>(smem + kChBiasOff);
    for (int i = threadIdx.x; i < jobs.n * kChUnits * 256; i += kChThreads)
      sb[i] = jobs.c[i >> 10].u[(i >> 8) & 3].bias[i & 255];
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      uint32_t wit = 0, n_in = 0;
      auto load_w = [&](const CUtensorMap* tm, int k) {          // this CTA's 128 of the 256 weight rows, K chunk k
        const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
        mbar_wait(&w_empty[s], ph ^ 1);
        if (rank == 0) mbar_expect_tx(&w_full[s], 2 * kStageBytes);
        tma_load_2d_2sm(smem + kChWOff + s * kStageBytes, tm, k * kChunkK, (int)rank * 128, &w_full[s], kEvictLast);
        ++wit;
      };
      auto load_in = [&](int item) {
        const ChainJob& J = jobs.c[chain_of(item)];
        const int g0 = kGroup * group_of(item);
        const int tile0 = g0 + 2 * (int)rank;
        const int n_valid = min(kGroup, jobs.n_tiles - g0);
        mbar_wait(in_empty, (n_in & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(in_full, (uint32_t)n_valid * kChSlab);
        for (int t = 0; t < 2; ++t)
          if (tile0 + t < jobs.n_tiles) tma_load_2d_2sm(smem + kChInOff + t * kChSlab, &J.tmIn, 0, (tile0 + t) * kTileM, in_full, kEvictFirst);
        ++n_in;
      };
      if (cid < n_items) load_in(cid);
      for (int item = cid; item < n_items; item += ncl) {
        const ChainJob& J = jobs.c[chain_of(item)];
        for (int u = 0; u < kChUnits; ++u) {
          for (int t = 0; t < 2; ++t)
            for (int k = 0; k < (u == 0 ? 1 : 4); ++k) load_w(&J.u[u].tmW, k);
          // the next item's input as soon as unit 0 of this item has consumed the buffer (in_empty): long before it is needed
          if (u == 1 && item + ncl < n_items) load_in(item + ncl);
        }
        if (kOut) {                                   // the output layer's hi/lo rows: this CTA's 8 of 16, 1 KB per K chunk
          const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
          mbar_wait(&w_empty[s], ph ^ 1);
          if (rank == 0) mbar_expect_tx(&w_full[s], 4 * 2048);
          for (int k = 0; k < 4; ++k)
            tma_load_2d_2sm(smem + kChWOff + s * kStageBytes + k * 1024, &J.tmWout, k * kChunkK, (int)rank * 8, &w_full[s], kEvictLast);
          ++wit;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = idesc_bf16(256, 256, 0, 0);
      constexpr uint32_t idesc_out = idesc_bf16(256, 16, 0, 0);
      uint32_t wit = 0, n_in = 0, c_epi[2] = {0u, 0u};
      bool first_item = true;
      for (int item = cid; item < n_items; item += ncl, first_item = false) {
        mbar_wait(in_full, n_in & 1);
        ++n_in;
        tc_fence_after();
        for (int u = 0; u < kChUnits; ++u) {
          for (uint32_t t = 0; t < 2; ++t) {
            if (!(first_item && u == 0)) {           // tile t's previous epilogue: slabs rewritten, accumulator drained
              mbar_wait(&epi_done[t], c_epi[t] & 1);
              ++c_epi[t];
              tc_fence_after();
            }
            const int nk = u == 0 ? 1 : 4;
            for (int k = 0; k < nk; ++k) {
              const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
              mbar_wait(&w_full[s], ph);
              tc_fence_after();
              const uint32_t a = u == 0 ? s_in + t * kChSlab : s_act + (t * 4 + (uint32_t)k) * kChSlab;
              const uint64_t da = smem_desc_sw128(a, 16, 1024);
              const uint64_t db = smem_desc_sw128(s_w + s * kStageBytes, 16, 1024);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(tmem_base + t * 256, da + 2 * j, db + 2 * j, idesc, (k | j) != 0);
              umma_commit_2sm(&w_empty[s]);
              ++wit;
            }
            umma_commit_2sm(&acc_full[t]);
          }
          if (u == 0) umma_commit_2sm(in_empty);
        }
        if (kOut) {
          // output layer: N = 16 (rows 0..7 hi, 8..15 lo halves of the <= 4 real output rows) into columns 0..15 of acc[tile]
          const uint32_t s = wit % kStages, ph = (wit / kStages) & 1;
          mbar_wait(&w_full[s], ph);
          tc_fence_after();
          for (uint32_t t = 0; t < 2; ++t) {
            mbar_wait(&epi_done[t], c_epi[t] & 1);
            ++c_epi[t];
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = smem_desc_sw128(s_act + (t * 4 + k) * kChSlab, 16, 1024);
              const uint64_t db = smem_desc_sw128(s_w + s * kStageBytes + k * 1024, 16, 1024);
#pragma unroll
              for (int j = 0; j < 4; ++j) umma_bf16_2sm(tmem_base + t * 256, da + 2 * j, db + 2 * j, idesc_out, (k | j) != 0);
            }
            umma_commit_2sm(&acc_full[t]);
          }
          umma_commit_2sm(&w_empty[s]);
          ++wit;
        }
      }
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue: group = tile of the CTA, thread = row
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const bool gleader = threadIdx.x == 64 + 128 * grp;
    const uint32_t sw_row = (uint32_t)(r >> 3) * 1024 + (uint32_t)(r & 7) * 128;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + grp * 256;
    uint32_t c_acc = 0;
    uint32_t* pend_flag = nullptr;       // (group leader) flag of the newest stored tile that is not published yet
    const uint32_t epi_bar = mapa_u32(smem_u32(&epi_done[grp]), 0);       // the leader CTA's epi_done[tile]
    for (int item = cid; item < n_items; item += ncl) {
      const int chain = chain_of(item);
      const ChainJob& J = jobs.c[chain];
      const int tile = kGroup * group_of(item) + 2 * (int)rank + grp;
      const bool valid = tile < jobs.n_tiles;
      const size_t grow = (size_t)tile * kTileM + r;
      for (int u = 0; u < kChUnits; ++u) {
        const ChainUnit& U = J.u[u];
        uint4 mb0 = make_uint4(0u, 0u, 0u, 0u), mb1 = mb0;
        if (MODE == CH_DX && valid) {
          mb0 = *reinterpret_cast<const uint4*>(U.bits + grow * J.bits_ld);
          mb1 = *reinterpret_cast<const uint4*>(U.bits + grow * J.bits_ld + 4);
        }
        mbar_wait(&acc_full[grp], c_acc & 1);
        ++c_acc;
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          if (valid) {
            // the TMA stores that read slabs 2h, 2h+1 (this group's second-newest commit group) must have finished reading
            if (gleader) bulk_wait_read<1>();
            named_bar_sync(1 + grp, 128);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int j = 2 * h + jj;
              uint32_t v[64];
              {
                uint32_t (&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
                uint32_t (&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
                tmem_ld32(t_lane + j * 64, v0);
                tmem_ld32(t_lane + j * 64 + 32, v1);
              }
              const uint32_t ob = s_act + (uint32_t)(grp * 4 + j) * kChSlab + sw_row;
              uint32_t obits[2] = {0u, 0u};
              const uint4 mbh = h == 0 ? mb0 : mb1;
              const uint32_t mw[2] = {jj == 0 ? mbh.x : mbh.z, jj == 0 ? mbh.y : mbh.w};
              const uint32_t bp = smem_u32(smem + kChBiasOff) + (uint32_t)((chain * kChUnits + u) * 256 + j * 64) * 4;
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                uint32_t w[4];
                float bv[8];
                if (MODE == CH_FWD) {
                  const float4 b0 = lds128f(bp + i * 32), b1 = lds128f(bp + i * 32 + 16);
                  bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
                }
#pragma unroll
                for (int pr = 0; pr < 4; ++pr) {
                  const int c = i * 8 + pr * 2;             // column inside the slab
                  const int tt = (c & 31) >> 1;             // pair index inside the 32-column group
                  float lo = __uint_as_float(v[c]), hi = __uint_as_float(v[c + 1]);
                  if (MODE == CH_FWD) {
                    lo += bv[pr * 2];
                    hi += bv[pr * 2 + 1];
                    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                    uint32_t gt;
                    asm("set.gt.u32.bf16x2 %0, %1, %2;" : "=r"(gt) : "r"(w[pr]), "r"(0u));
                    obits[c >> 5] |= gt & (0x00010001u << tt);
                  } else {
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[pr]) : "f"(hi), "f"(lo));
                    const uint32_t sel = (mw[c >> 5] >> tt) & 0x00010001u;
                    w[pr] &= sel * 0xFFFFu;
                  }
                }
                sts128(ob + (((uint32_t)i ^ (uint32_t)(r & 7)) << 4), w[0], w[1], w[2], w[3]);
              }
              if (MODE == CH_FWD)
                *reinterpret_cast<uint2*>(U.bits + grow * J.bits_ld + j * 2) = make_uint2(obits[0], obits[1]);
            }
            fence_proxy_async_smem();
            named_bar_sync(3 + grp, 128);                  // all 128 rows of both slabs are in SMEM
            if (gleader) {
              const uint64_t pol = MODE == CH_FWD ? kEvictFirst : (jobs.store_last ? kEvictLast : kEvictNormal);
              tma_store_2d_hint(&U.tmOut, (2 * h) * 64, tile * kTileM, smem + kChActOff + (grp * 4 + 2 * h) * kChSlab, pol);
              tma_store_2d_hint(&U.tmOut, (2 * h + 1) * 64, tile * kTileM, smem + kChActOff + (grp * 4 + 2 * h + 1) * kChSlab, pol);
              bulk_commit();
              if (MODE == CH_DX && jobs.ready && h == 1) {
                // lagged hand-over to the dW pairs (see chain_role)
                if (pend_flag) {
                  bulk_wait<2>();
                  fence_proxy_async_all();
                  st_release_gpu(pend_flag, jobs.epoch);
                }
                pend_flag = jobs.ready + ((size_t)(chain * kChUnits + u) * jobs.n_tiles + tile);
              }
            }
          }
        }
        // the MMA warp may read the tile's slabs / overwrite its accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(epi_bar);
      }
      if (kOut) {
        mbar_wait(&acc_full[grp], c_acc & 1);
        ++c_acc;
        tc_fence_after();
        if (valid) {
          uint32_t v[16];
          tmem_ld16(t_lane, v);
          tmem_ld_wait();
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = e < J.k_out ? __uint_as_float(v[e]) + __uint_as_float(v[8 + e]) + __ldg(J.bias_out + e) : 0.f;
          *reinterpret_cast<float4*>(J.logits + grow * 4) = make_float4(o[0], o[1], o[2], o[3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(epi_bar);
      }
    }
    if (gleader) {
      bulk_wait<0>();
      if (MODE == CH_DX && pend_flag) {
        fence_proxy_async_all();
        st_release_gpu(pend_flag, jobs.epoch);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();        // the peer may still signal this CTA's barriers / read its SMEM until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// SCHED 0: two 128-column halves per layer, both tiles in lockstep (chain_role); SCHED 1: tile-staggered N = 256 (CTA pairs only)
template <int MODE, int CL, int SCHED = 0>
__global__ void __launch_bounds__(kChThreads, 1) k_tc_chain(const __grid_constant__ ChainJobs jobs) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (SCHED == 1 && CL == 2) chain_role_staggered<MODE>(jobs, smem, (int)blockIdx.x / CL, (int)gridDim.x / CL);
  else chain_role<MODE, CL>(jobs, smem, (int)blockIdx.x / CL, (int)gridDim.x / CL);
}

}  // namespace tc
}  // namespace marf
