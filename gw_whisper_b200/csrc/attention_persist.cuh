// Persistent variant of the fused attention kernel (attention_tc.cuh has the algorithm notes and the
// reference citations: HF modeling_whisper.py:215-238, q pre-scaled through W_q).
//
// Why: a work item (det-window, head, pair of 128-row query tiles) is only 12 key tiles long at
// T = 1500.  With one CTA per item, the CTA's prologue (barrier init, TMEM alloc, Q and first K/V
// loads at full L2/HBM latency) and epilogue (O read-out, store, dealloc, CTA turnover) were exposed:
// the r1 T-sweep measured ~3 us of fixed cost per 24.7 us CTA plus a slower first / last tile, while
// the steady state already runs near the instruction-mix ceiling of the softmax (tools/ubench/
// softmax_mix.cu: 12.7 elements/clk/SM with two warps per sub-partition, because F2FP shares the XU
// pipe with MUFU.EX2).  Here one CTA per SM loops over items:
//   * the TMA producer streams K/V tiles of consecutive items through the same ring and prefetches the
//     next item's Q pair into the other half of a double buffer,
//   * the MMA warp issues S(0) of the next item as soon as the softmax group has copied the last S of
//     the current item out of TMEM, i.e. while that group is still in its last tile / epilogue,
//   * barriers, TMEM and tensor maps are set up once per SM.
// Everything else (TMEM map, softmax, lazy rescale, event-driven issue, P in TMEM) is as in
// attention_tc_kernel<2>.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace gww {

struct AttnPersistParams {
  int T;          // tokens per det-window (1500)
  int d_model;    // 384 / 512 / 768
  int nkv;        // ceil(T / 128)
  int n_heads;    // d_model / 64
  int n_qpairs;   // ceil(T / 256)
  int n_items;    // n_qpairs * n_heads * det-windows
};

#ifndef GWW_AP_STAGES
#define GWW_AP_STAGES 3
#endif
constexpr int kApStages = GWW_AP_STAGES;   // K/V ring depth (each stage = one 16 KB K tile + one 16 KB V tile)
// Exponentials are issued in groups of kApGroup; the arguments of group g depend (through a
// multiply-by-zero FFMA) on the row-sum accumulator as it stands after group g - kApLookahead, so at most
// kApGroup * kApLookahead MUFUs of a warp are queued at a time: the MIO queue of the sub-partition, which
// the MMA / TMA warps' mbarrier instructions also go through, is not permanently full of MUFUs.  0 = off.
#ifndef GWW_AP_GROUP
#define GWW_AP_GROUP 16    // r2 sweep with the polynomial share below (profiles/r2_attn_tune.jsonl): 4 -> 16
#endif
#ifndef GWW_AP_LOOKAHEAD
#define GWW_AP_LOOKAHEAD 2
#endif
constexpr int kApGroup = GWW_AP_GROUP;
constexpr int kApLookahead = GWW_AP_LOOKAHEAD;
#ifndef GWW_AP_PACKED_F32
#define GWW_AP_PACKED_F32 1   // scale-subtract and row-sum on packed fp32 pairs (FFMA2 / FADD2)
#endif
// Of every 8 consecutive pairs of scores, GWW_AP_POLY pairs take their exponentials from the FMA pipe
// (poly_exp2_pair) instead of MUFU.EX2: the kernel is bound by the XU pipe (MUFU.EX2 16 / clk / SM, r1 ncu: XU 68 %,
// tensor 34 %), the fp32 pipe has slack.  Not used in the masked last key tile (-inf scores).  0 = all MUFU.
// r2 sweep, 256 det-windows of whisper-base, ms per launch: (poly, group, lookahead) = (0,4,2) 1.654 [round 1],
// (0,0,2) 1.625, (2,0,2) 1.625, (2,8,2) 1.566, (2,16,2) 1.562, (3,0,2) 1.572, (4,0,2) 1.707, (4,8,3) 1.673.
#ifndef GWW_AP_POLY
#define GWW_AP_POLY 2
#endif
// GWW_AP_MMA_WARPS = 2: one MMA-issuing warp per query tile (warps 9 and 10, two different sub-partitions), each a
// static sequence S(0), [S(m), P.V(m-1)]..., P.V(last) on blocking mbarrier waits -- within one query tile "S copied
// out" always precedes "P written", so the static order is the event order.  Why (profiles/r2_attn_trace.txt): the
// softmax warps' MUFUs go through the same in-order MIO queue of a sub-partition as the issuer's mbarrier and tcgen05
// instructions; a ready mbarrier test costs the issuer 70-230 clk and one block of 4-8 tcgen05.mma + commits ~450 clk
// (measured by letting the softmax warps issue: -DGWW_AP_TRACE variants of this round).  A single event-driven
// issuer therefore needs ~600 clk per event x 4 events (S0, S1, PV0, PV1) of every ~2870-clk tile period: it is
// ~85 % busy, and "P written" -> "P.V finished" took 1000-1900 clk for 256 clk of tensor work.  1 = that single
// event-driven issuer (round 1).
#ifndef GWW_AP_MMA_WARPS
#define GWW_AP_MMA_WARPS 2
#endif
#ifndef GWW_AP_EARLY_TEST
#define GWW_AP_EARLY_TEST 1   // softmax warps test sfull(n+1) / pvdone(n-1) ahead of use (non-blocking)
#endif
#ifndef GWW_ATTN_POLL_NS
#define GWW_ATTN_POLL_NS 64   // back-off of the MMA warp's polling loop when nothing is ready
#endif
// GWW_ATTN_TRACE (tuning builds only): the first CTAs print where warp 0 of softmax group 0 spent its
// clocks (barrier waits vs work) and the effective SM clock (clock64 against globaltimer).
#ifdef GWW_ATTN_TRACE
#define AP_TRACE(x) x
#else
#define AP_TRACE(x)
#endif
// Q double buffer (2 x 2 tiles x 16 KB) | K ring | V ring | O staging (2 tiles x 16 KB) | barriers
constexpr int kApSmemBytes = 65536 + kApStages * 2 * 16384 + 32768 + 512;

__global__ void __launch_bounds__(384, 1)
attention_persist_kernel(const __grid_constant__ CUtensorMap tmQKV,  // {3d, T, Bt} box {64,128,1}
                         const __grid_constant__ CUtensorMap tmO,    // {d, T, Bt}  box {64,32,1}
                         const AttnPersistParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("gww: attention dynamic smem base not 1024-aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  uint8_t* q_s = smem;                                  // [2 buffers][2 tiles] x 16 KB
  uint8_t* k_s = smem + 65536;                          // kApStages x 16 KB
  uint8_t* v_s = k_s + kApStages * 16384;               // kApStages x 16 KB
  uint8_t* o_s = v_s + kApStages * 16384;               // [2 tiles][4 warps] x 4 KB output staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(o_s + 32768);
  const uint32_t bar_qfull = smem_u32(bars);                // 2
  const uint32_t bar_qempty = bar_qfull + 16;               // 2
  const uint32_t bar_kfull = bar_qempty + 16;               // kApStages
  const uint32_t bar_kempty = bar_kfull + 8 * kApStages;
  const uint32_t bar_vfull = bar_kempty + 8 * kApStages;
  const uint32_t bar_vempty = bar_vfull + 8 * kApStages;
  const uint32_t bar_sfull = bar_vempty + 8 * kApStages;    // 2 (mma -> softmax: S_t ready)
  const uint32_t bar_sfree = bar_sfull + 16;                // 2 (softmax -> mma: S_t copied to registers)
  // (a variant that split P_t.V into two 64-key halves, to overlap the first half's MMAs with the second
  //  half's exponentials, was measured slower: 1.95 vs 1.74 ms -- the extra barrier traffic costs the MMA
  //  warp more than the overlap gains)
  const uint32_t bar_pfull = bar_sfree + 16;                // 2 (softmax -> mma: P_t written)
  const uint32_t bar_pvdone = bar_pfull + 16;               // 2 (mma -> softmax: P_t.V finished)
  const uint32_t bar_ofull = bar_pvdone + 16;               // 2 (mma -> softmax: O_t of the item complete)
  const uint32_t bar_ofree = bar_ofull + 16;                // 2 (softmax -> mma: O_t read out)
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 4 + 4 * kApStages + 12);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int nkv = p.nkv;
  // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  (consecutive CTAs work on the query
  // pairs of the same (head, det-window) at the same time: its K/V stay hot in L2)
  const int n_local = (static_cast<int>(blockIdx.x) < p.n_items)
                          ? (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                          : 0;
  auto item_coords = [&](int k, int& q0, int& head, int& bi) {
    const int item = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
    const int qh = p.n_qpairs * p.n_heads;
    bi = item / qh;
    const int r = item - bi * qh;
    head = r / p.n_qpairs;
    q0 = (r - head * p.n_qpairs) * 256;
  };

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_qfull + 8 * i, 1);
      mbar_init(bar_qempty + 8 * i, GWW_AP_MMA_WARPS);        // one commit per issuing warp
      mbar_init(bar_sfull + 8 * i, 1);
      mbar_init(bar_sfree + 8 * i, 4);       // one elected arrival per softmax warp
      mbar_init(bar_pfull + 8 * i, 4);
      mbar_init(bar_pvdone + 8 * i, 1);
      mbar_init(bar_ofull + 8 * i, 1);
      mbar_init(bar_ofree + 8 * i, 4);
    }
    for (int i = 0; i < kApStages; ++i) {
      mbar_init(bar_kfull + 8 * i, 1);
      mbar_init(bar_kempty + 8 * i, GWW_AP_MMA_WARPS);
      mbar_init(bar_vfull + 8 * i, 1);
      mbar_init(bar_vempty + 8 * i, GWW_AP_MMA_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc<512>(smem_u32(tmem_ptr_s));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (wg == 2) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (warp == 8) {
      // ===================== TMA producer =====================
      const uint32_t q_smem = smem_u32(q_s), k_smem = smem_u32(k_s), v_smem = smem_u32(v_s);
      int stage = 0;
      uint32_t phase = 0;
      for (int k = 0; k < n_local; ++k) {
        int q0, head, bi;
        item_coords(k, q0, head, bi);
        const int qb = k & 1;
        mbar_wait(bar_qempty + 8 * qb, ((k >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_qfull + 8 * qb, 32768);
          tma_load_3d(q_smem + qb * 32768, &tmQKV, bar_qfull + 8 * qb, head * 64, q0, bi);
          tma_load_3d(q_smem + qb * 32768 + 16384, &tmQKV, bar_qfull + 8 * qb, head * 64, q0 + 128, bi);
        }
        __syncwarp();
        const int kcol = p.d_model + head * 64, vcol = 2 * p.d_model + head * 64;
        for (int j = 0; j < nkv; ++j) {
          mbar_wait(bar_kempty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_kfull + 8 * stage, 16384);
            tma_load_3d(k_smem + stage * 16384, &tmQKV, bar_kfull + 8 * stage, kcol, j * 128, bi);
          }
          __syncwarp();
          mbar_wait(bar_vempty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_vfull + 8 * stage, 16384);
            tma_load_3d(v_smem + stage * 16384, &tmQKV, bar_vfull + 8 * stage, vcol, j * 128, bi);
          }
          __syncwarp();
          if (++stage == kApStages) { stage = 0; phase ^= 1; }
        }
      }
#if GWW_AP_MMA_WARPS == 2
    } else if (warp == 9 || warp == 10) {
      // ===================== MMA issuer of query tile t: static order, blocking waits =====================
      const int t = warp - 9;
      constexpr uint32_t kIdescS = make_idesc_op16(128, 128, 0);
      constexpr uint32_t kIdescO = make_idesc_op16(128, 64, 1);   // V is MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      const uint32_t tS = tb + t * 128, tP = tb + 256 + t * 64, tO = tb + 384 + t * 64;
      const uint32_t q_smem = smem_u32(q_s) + t * 16384, k_smem = smem_u32(k_s), v_smem = smem_u32(v_s);
      const uint32_t b_sfull = bar_sfull + 8 * t, b_sfree = bar_sfree + 8 * t;
      const uint32_t b_pfull = bar_pfull + 8 * t, b_pvdone = bar_pvdone + 8 * t;
      const int total = n_local * nkv;
      int s_k = 0, s_j = 0, s_stage = 0, pv_k = 0, pv_j = 0, pv_stage = 0;
      uint32_t s_phase = 0, pv_phase = 0;
      for (int m = 0; m <= total; ++m) {
        if (m < total) {
          // ---- S(m) = Q.K^T of tile m: K tile resident, S(m-1) copied out, (first tile of an item) Q resident
          // (the barrier that completes last is waited for last: every wait, even on a completed barrier, is a
          //  round trip through the MIO queue behind the softmax warps' MUFUs)
          if (s_j == 0) mbar_wait(bar_qfull + 8 * (s_k & 1), (s_k >> 1) & 1);
          if (m > 0) mbar_wait2(bar_kfull + 8 * s_stage, s_phase, b_sfree, (m - 1) & 1);
          else mbar_wait(bar_kfull + 8 * s_stage, s_phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t qdesc = make_sw128_desc(q_smem + (s_k & 1) * 32768);
            const uint64_t kdesc = make_sw128_desc(k_smem + s_stage * 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(tS, qdesc + 2 * kk, kdesc + 2 * kk, kIdescS, kk);
            umma_commit(b_sfull);
            umma_commit(bar_kempty + 8 * s_stage);                    // second arrival: the other tile's issuer
            if (s_j + 1 == nkv) umma_commit(bar_qempty + 8 * (s_k & 1));
          }
          __syncwarp();
          if (++s_j == nkv) { s_j = 0; ++s_k; }
          if (++s_stage == kApStages) { s_stage = 0; s_phase ^= 1; }
        }
        if (m > 0) {
          // ---- O += P(m-1).V(m-1): V tile resident, P written, (first tile of an item) previous O read out
          if (pv_j == 0 && pv_k > 0) mbar_wait(bar_ofree + 8 * t, (pv_k - 1) & 1);
          mbar_wait2(bar_vfull + 8 * pv_stage, pv_phase, b_pfull, (m - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t vdesc = make_sw128_desc(v_smem + pv_stage * 16384);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_ts(tO, tP + 8 * kk, vdesc + 128 * kk, kIdescO, (pv_j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(b_pvdone);
            umma_commit(bar_vempty + 8 * pv_stage);
            if (pv_j + 1 == nkv) umma_commit(bar_ofull + 8 * t);
          }
          __syncwarp();
          if (++pv_j == nkv) { pv_j = 0; ++pv_k; }
          if (++pv_stage == kApStages) { pv_stage = 0; pv_phase ^= 1; }
        }
      }
#else
    } else if (warp == 9) {
      // ===================== MMA issuer (event driven, see attention_tc.cuh) =====================
      constexpr uint32_t kIdescS = make_idesc_op16(128, 128, 0);
      constexpr uint32_t kIdescO = make_idesc_op16(128, 64, 1);   // V is MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      const uint32_t tS[2] = {tb + 0u, tb + 128u};
      const uint32_t tP[2] = {tb + 256u, tb + 320u};
      const uint32_t tO[2] = {tb + 384u, tb + 448u};
      const uint32_t q_smem = smem_u32(q_s), k_smem = smem_u32(k_s), v_smem = smem_u32(v_s);
      const int total = n_local * nkv;          // key tiles per query tile over all items of this CTA
      // per query tile t: next S / next P.V in the global tile sequence, decomposed as (item k, tile j)
      int s_seq[2] = {0, 0}, s_k[2] = {0, 0}, s_j[2] = {0, 0}, s_stage[2] = {0, 0};
      int pv_seq[2] = {0, 0}, pv_k[2] = {0, 0}, pv_j[2] = {0, 0}, pv_stage[2] = {0, 0};
      uint32_t s_phase[2] = {0, 0}, pv_phase[2] = {0, 0};
      uint64_t idle_t0 = 0;
      while (pv_seq[0] < total || pv_seq[1] < total) {
        bool progressed = false;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (s_seq[t] < total) {
            const int n = s_seq[t], k = s_k[t], j = s_j[t], st = s_stage[t];
            uint32_t ok = mbar_test(bar_kfull + 8 * st, s_phase[t]);
            if (n > 0) ok &= mbar_test(bar_sfree + 8 * t, (n - 1) & 1);            // S_t(n-1) copied out
            if (j == 0) ok &= mbar_test(bar_qfull + 8 * (k & 1), (k >> 1) & 1);    // Q pair of item k resident
            if (__all_sync(0xffffffffu, ok)) {
              tc_fence_after();
              if (elect_one()) {
                const uint64_t qdesc = make_sw128_desc(q_smem + (k & 1) * 32768 + t * 16384);
                const uint64_t kdesc = make_sw128_desc(k_smem + st * 16384);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_ss(tS[t], qdesc + 2 * kk, kdesc + 2 * kk, kIdescS, kk);
                umma_commit(bar_sfull + 8 * t);
                if (s_seq[t ^ 1] > n) {                    // both query tiles have issued on K tile n
                  umma_commit(bar_kempty + 8 * st);
                  if (j + 1 == nkv) umma_commit(bar_qempty + 8 * (k & 1));   // ... and are done with item k's Q
                }
              }
              __syncwarp();
              s_seq[t] = n + 1;
              if (++s_j[t] == nkv) { s_j[t] = 0; ++s_k[t]; }
              if (++s_stage[t] == kApStages) { s_stage[t] = 0; s_phase[t] ^= 1; }
              progressed = true;
            }
          }
          if (pv_seq[t] < total) {
            const int n = pv_seq[t], k = pv_k[t], j = pv_j[t], st = pv_stage[t];
            uint32_t ok = mbar_test(bar_vfull + 8 * st, pv_phase[t]) & mbar_test(bar_pfull + 8 * t, n & 1);
            if (j == 0 && k > 0) ok &= mbar_test(bar_ofree + 8 * t, (k - 1) & 1);  // previous item's O read out
            if (__all_sync(0xffffffffu, ok)) {
              tc_fence_after();
              if (elect_one()) {
                const uint64_t vdesc = make_sw128_desc(v_smem + st * 16384);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
                  umma_ts(tO[t], tP[t] + 8 * kk, vdesc + 128 * kk, kIdescO, (j > 0 || kk > 0) ? 1u : 0u);
                umma_commit(bar_pvdone + 8 * t);
                if (pv_seq[t ^ 1] > n) umma_commit(bar_vempty + 8 * st);   // both query tiles have issued on V tile n
                if (j + 1 == nkv) umma_commit(bar_ofull + 8 * t);
              }
              __syncwarp();
              pv_seq[t] = n + 1;
              if (++pv_j[t] == nkv) { pv_j[t] = 0; ++pv_k[t]; }
              if (++pv_stage[t] == kApStages) { pv_stage[t] = 0; pv_phase[t] ^= 1; }
              progressed = true;
            }
          }
        }
        if (progressed) {
          idle_t0 = 0;
        } else {
          nanosleep_ns(GWW_ATTN_POLL_NS);
          const uint64_t now = globaltimer_ns();
          if (idle_t0 == 0) idle_t0 = now;
          else if (now - idle_t0 > GWW_MBAR_TIMEOUT_NS) {
            if (lane == 0)
              printf("gww: persistent attention issue loop timeout block=%d s_seq=(%d,%d) pv_seq=(%d,%d) total=%d\n",
                     blockIdx.x, s_seq[0], s_seq[1], pv_seq[0], pv_seq[1], total);
            __trap();
          }
        }
      }
#endif
    }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int t = wg;                           // query tile handled by this warpgroup
    const int wq = warp & 3;                    // TMEM lane quarter
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tP = tmem_base + lane_off + 256 + t * 64;
    const uint32_t tO = tmem_base + lane_off + 384 + t * 64;
    constexpr float kLog2e = 1.4426950408889634f;
    float m_used = 0.f, l = 0.f;
    const uint32_t b_sfull = bar_sfull + 8 * t, b_sfree = bar_sfree + 8 * t;
    const uint32_t b_pfull = bar_pfull + 8 * t, b_pvdone = bar_pvdone + 8 * t;
    int n = 0;                                  // key tiles processed so far (barrier parities)
    bool s_early = false;                       // "S of tile n is ready" already observed
    AP_TRACE(long long w_sfull = 0; long long w_pv = 0; long long w_ofull = 0; long long w_ld = 0; long long w_exp = 0;)
    AP_TRACE(long long w_max = 0; long long w_tail = 0; long long w_epi = 0;)
    AP_TRACE(const long long c_begin = clock64(); const unsigned long long ns_begin = globaltimer_ns();)

    // one 128-key tile: j = tile within the item (controls masking / first-tile handling)
    auto tile = [&](const int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      AP_TRACE(long long c0 = clock64();)
      if (!s_early) mbar_wait(b_sfull, n & 1);     // usually already seen complete by the test inside the previous tile
      AP_TRACE(long long c1 = clock64(); w_sfull += c1 - c0;)
      tc_fence_after();
      uint32_t s[4][32];
      tmem_ld32(tS + 0, s[0]);
      tmem_ld32(tS + 32, s[1]);
      tmem_ld32(tS + 64, s[2]);
      tmem_ld32(tS + 96, s[3]);
      tmem_wait_ld();
      AP_TRACE(long long c1b = clock64(); w_ld += c1b - c1;)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_sfree);       // S_t is in registers: the next Q.K^T may overwrite it
      if constexpr (kMask) {
        const int valid = p.T - j * 128;         // keys [0, valid) of this tile exist
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
      }
      float mt0 = -INFINITY, mt1 = -INFINITY, mt2 = -INFINITY, mt3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mt0 = fmaxf(mt0, __uint_as_float(s[0][i]));
        mt1 = fmaxf(mt1, __uint_as_float(s[1][i]));
        mt2 = fmaxf(mt2, __uint_as_float(s[2][i]));
        mt3 = fmaxf(mt3, __uint_as_float(s[3][i]));
      }
      const float mt = fmaxf(fmaxf(mt0, mt1), fmaxf(mt2, mt3));
      // at j == 0 the previous item's O / P are already free: its epilogue waited for o_full
      bool pv_waited = (j == 0);
      if (j == 0) {
        m_used = mt;
      } else {
        // lazy rescale: exact (same algebra as online softmax), skipped while the stale max keeps
        // exp2 arguments <= 8
        const bool need = (mt - m_used) * kLog2e > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(b_pvdone, (n - 1) & 1);      // O_t must be complete before it is rescaled
          tc_fence_after();
          pv_waited = true;
          const float sc = need ? fast_exp2((m_used - mt) * kLog2e) : 1.0f;
          uint32_t o[16];
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            tmem_ld16(tO + h * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
            tmem_st16(tO + h * 16, o);
          }
          tmem_wait_st();
          l *= sc;
          if (need) m_used = mt;
        }
      }
      // non-blocking look at "P(n-1).V finished" after the first 32 exponentials, consumed after 64
      uint32_t pv_early = 0u;
      const float mneg = -m_used * kLog2e;
      float l0 = 0.f, l1 = 0.f;
      uint32_t pk[32];
      constexpr int kGroups = (kApGroup > 0) ? 128 / kApGroup : 1;
      float lsnap[kGroups];
      float mneg_g = mneg;
      AP_TRACE(asm volatile("" ::"f"(mneg));)
      AP_TRACE(long long c2 = clock64(); w_max += c2 - c1b;)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if constexpr (kApGroup > 0) {
            const int e = c * 32 + i;
            if (e % (kApGroup > 0 ? kApGroup : 1) == 0) {
              const int g = e / (kApGroup > 0 ? kApGroup : 1);
              if (g > 0) lsnap[g - 1] = l0;
              mneg_g = (g >= kApLookahead) ? fmaf(lsnap[g - kApLookahead], 0.0f, mneg) : mneg;   // == mneg
            }
          }
#if GWW_AP_PACKED_F32
          float x0, x1;                              // FFMA2 / FADD2: half the fp32-pipe instructions
          ffma2_bcast(x0, x1, __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]), kLog2e, mneg_g);
          float p0, p1;
          if (!kMask && GWW_AP_POLY > 0 && ((i >> 1) & 7) < GWW_AP_POLY) {
            poly_exp2_pair(f2_pack(fmaxf(x0, -125.f), fmaxf(x1, -125.f)), p0, p1);
          } else {
            p0 = fast_exp2(x0);
            p1 = fast_exp2(x1);
          }
          fadd2_acc(l0, l1, p0, p1);
#else
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[c][i]), kLog2e, mneg_g));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[c][i + 1]), kLog2e, mneg_g));
          l0 += p0;
          l1 += p1;
#endif
          pk[(c & 1) * 16 + (i >> 1)] = pack_op16x2(p0, p1);
        }
        if (GWW_AP_EARLY_TEST && c == 0) pv_early = mbar_test(b_pvdone, (n - 1) & 1);   // unconditional: result unused at j == 0
        if (GWW_AP_EARLY_TEST && c == 2) s_early = __all_sync(0xffffffffu, mbar_test(b_sfull, (n + 1) & 1));   // S of the next tile
        if (c & 1) {                                 // 64 keys packed -> 32 TMEM columns of P: half c >> 1
          if (!pv_waited) {
            AP_TRACE(long long c3 = clock64();)
            if (!__all_sync(0xffffffffu, pv_early)) mbar_wait(b_pvdone, (n - 1) & 1);   // P_t(n-1).V finished reading P_t
            AP_TRACE(long long c4 = clock64(); w_pv += c4 - c3; c2 += c4 - c3;)
            tc_fence_after();
            pv_waited = true;
          }
          tmem_st32(tP + (c >> 1) * 32, pk);
        }
      }
      l += l0 + l1;
      AP_TRACE(asm volatile("" ::"f"(l), "r"(pk[31]));)
      AP_TRACE(long long c7 = clock64(); w_exp += c7 - c2;)
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
      AP_TRACE(w_tail += clock64() - c7;)
      ++n;
    };

    uint8_t* stg = o_s + t * 16384 + wq * 4096;       // this warp's 32-row x 128-byte output box
    uint8_t* sb = stg + lane * 128;
    for (int k = 0; k < n_local; ++k) {
      int q0, head, bi;
      item_coords(k, q0, head, bi);
      l = 0.f;
      for (int j = 0; j + 1 < nkv; ++j) tile(j, std::false_type{});
      if ((p.T & 127) != 0) tile(nkv - 1, std::true_type{});
      else tile(nkv - 1, std::false_type{});
      // ---- item epilogue: O_t / l -> bf16 -> swizzled staging -> TMA store
      AP_TRACE(long long c5 = clock64();)
      mbar_wait(bar_ofull + 8 * t, k & 1);
      AP_TRACE(long long c5b = clock64(); w_ofull += c5b - c5;)
      tc_fence_after();
      const float inv_l = 1.0f / l;
      uint32_t o0[32], o1[32];
      tmem_ld32(tO, o0);
      tmem_ld32(tO + 32, o1);
      tmem_wait_ld();
      tc_fence_before();
      if (lane == 0) tma_store_wait_read<0>();    // this warp's previous store has drained the staging box
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ofree + 8 * t);   // O_t read out: the next item's first P.V may overwrite it
#pragma unroll
      for (int jc = 0; jc < 8; ++jc) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = jc * 8 + e * 2;
          const float a0 = __uint_as_float(col < 32 ? o0[col & 31] : o1[col & 31]) * inv_l;
          const float a1 = __uint_as_float(col < 32 ? o0[(col + 1) & 31] : o1[(col + 1) & 31]) * inv_l;
          pk[e] = pack_op16x2(a0, a1);
        }
        *reinterpret_cast<uint4*>(sb + ((jc ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmO, smem_u32(stg), head * 64, q0 + t * 128 + wq * 32, bi);
        tma_store_commit();
      }
      AP_TRACE(w_epi += clock64() - c5b;)
    }
    if (lane == 0) tma_store_wait_all<0>();
    AP_TRACE(if (blockIdx.x < 1 && lane == 0) {
      const long long cyc = clock64() - c_begin;
      const unsigned long long ns = globaltimer_ns() - ns_begin;
      printf("gww-trace block=%d warp=%d tiles=%d cycles=%lld ns=%llu MHz=%.0f per-tile: total=%.0f sfull_wait=%.0f tmem_ld=%.0f max+rescale=%.0f exp_phase=%.0f (pvdone_wait=%.0f) tail=%.0f | per item: ofull_wait=%.0f epilogue=%.0f\n",
             blockIdx.x, warp, n, cyc, ns, 1e3 * (double)cyc / (double)ns, (double)cyc / n, (double)w_sfull / n, (double)w_ld / n,
             (double)w_max / n, (double)w_exp / n, (double)w_pv / n, (double)w_tail / n, (double)w_ofull / n_local, (double)w_epi / n_local);
    })
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace gww
