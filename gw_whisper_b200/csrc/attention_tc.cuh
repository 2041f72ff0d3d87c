// Fused non-causal multi-head self-attention for the Whisper encoder on tcgen05 / TMEM (sm_100a).
// Reference math: HF modeling_whisper.py:215-238 (eager_attention_forward, scaling=1.0, no mask,
// no dropout) with q pre-scaled by head_dim^-0.5 (:310) -- the scale is folded into W_q / b_q at
// model-load time, so this kernel computes softmax(Q K^T) V per (det-window, head).
//
// Layout: qkv [Bt, T, 3*d] bf16 (q | k | v, head h at columns h*64), out [Bt, T, d] bf16.
// One CTA = one (det-window, head, pair of 128-row query tiles); it streams the T keys in 128-row
// K/V tiles through a 3-stage TMA ring.
//   S_t = Q_t K_j^T     tcgen05.mma SS, M=128 N=128 K=64, accumulator in TMEM (fp32)
//   softmax             two warpgroups (one per query tile), one thread per row, online max with
//                       lazy rescale (only when the running max grows by > 2^8), exp2 on MUFU
//   P_t -> TMEM (bf16)  tcgen05.st, then O_t += P_t V_j as tcgen05.mma TS (A from TMEM, V is the
//                       MN-major B operand straight from its TMA tile), M=128 N=64 K=128
// TMEM map (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
// Warps: WG0/WG1 = softmax for query tile 0/1, WG2 = {w8: TMA producer, w9: MMA issuer + TMEM alloc}.
//
// Scheduling (r1 measurements: MUFU.EX2 is the binding pipe at 16 results/clk/SM; one softmax warp per
// SM sub-partition reaches only ~80 % of that, two reach 100 % -- profiles/r1_ubench_pipes.txt):
//   * S_t(j+1) is issued as soon as softmax t has copied S_t(j) into registers (s_free barrier): the
//     tensor work runs a whole tile ahead of the exponentials and is completely hidden;
//   * both softmax warpgroups run their exponentials concurrently (two warps per sub-partition keep
//     the MUFU saturated); no turn taking;
//   * lazy rescale: O and l are rescaled only when the tile max exceeds the running max by more
//     than 2^8 (rare); exact online-softmax algebra either way;
//   * kAttnFmaExp of every 32 exponentials are evaluated on the FMA/ALU pipes (round-to-nearest range
//     reduction + cubic, |rel err| < 7.5e-5, far below bf16 P's 2^-9) to take load off the MUFU;
//   * P_t(j).V(j) is issued when P is ready; softmax only waits for its completion (pv_done) right
//     before it overwrites P / rescales O for tile j+1; key masking exists only in the last tile's code.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace gww {

struct AttnParams {
  int T;        // tokens per det-window (1500)
  int d_model;  // 384 / 512 / 768
  int nkv;      // ceil(T / 128)
};

constexpr int kAttnStages = 3;
#ifndef GWW_ATTN_FMA_EXP
#define GWW_ATTN_FMA_EXP 0
#endif
constexpr int kAttnFmaExp = GWW_ATTN_FMA_EXP;   // exponentials per 32 evaluated without the MUFU
#ifndef GWW_ATTN_STAGGER
#define GWW_ATTN_STAGGER 0
#endif
// 1: warpgroup 1 starts its first tile half a tile behind warpgroup 0, so one group's TMEM loads /
// max / waits fall into the other group's exponentials instead of coinciding with them
constexpr int kAttnStagger = GWW_ATTN_STAGGER;
// Ablation builds for bottleneck hunting (WRONG numerics, tools/attn_bench.py only; default 0):
//   bit 0: odd exponentials skip the MUFU   bit 1: all exponentials skip the MUFU
//   bit 2: no row-sum FADDs                 bit 3: no row-max pass
//   bit 4: softmax never waits for s_full / pv_done (races)   bit 5: no TMEM loads / stores in softmax
#ifndef GWW_ATTN_ABLATE
#define GWW_ATTN_ABLATE 0
#endif
constexpr int kAttnAblate = GWW_ATTN_ABLATE;
// Exponentials are issued in groups of kAttnGroup; the arguments of group g are made to depend on the
// row-sum accumulator as it stands after group g - kAttnLookahead (a multiply-by-zero FFMA, one per
// group), so ptxas cannot hoist all 128 MUFUs of a tile into one burst ahead of their FADD / F2FP
// consumers: the MUFU and FMA pipes then overlap inside a warp instead of taking turns.  0 = off.
#ifndef GWW_ATTN_GROUP
#define GWW_ATTN_GROUP 8
#endif
#ifndef GWW_ATTN_LOOKAHEAD
#define GWW_ATTN_LOOKAHEAD 2
#endif
constexpr int kAttnGroup = GWW_ATTN_GROUP;
constexpr int kAttnLookahead = GWW_ATTN_LOOKAHEAD;
// NT = query tiles (of 128 rows) per CTA.  NT = 2 is the original shape (one 384-thread CTA per SM, both
// softmax groups fed by one producer / one MMA warp).  NT = 1 runs TWO independent 256-thread CTAs per
// SM (256 TMEM columns and 112 KB of shared memory each): a CTA lives for only 12 key tiles, and its
// prologue (barrier init, TMEM alloc, Q / first K load) and epilogue (O store, dealloc) -- measured at
// ~6 us of the 24.7 us a CTA took (T-sweep, r1) -- are then hidden behind the other CTA's steady state.
template <int NT>
struct AttnCfg {
  static constexpr int kThreads = NT * 128 + 128;
  static constexpr int kTmemCols = NT * 256;
  static constexpr int kQBytes = NT * 16384;
  static constexpr int kSmemBytes = kQBytes + kAttnStages * 2 * 16384 + 256;
  static constexpr int kCtasPerSm = (NT == 1) ? 2 : 1;
};
constexpr int kAttnSmemBytes = AttnCfg<2>::kSmemBytes;

template <int NT>
__global__ void __launch_bounds__(AttnCfg<NT>::kThreads, AttnCfg<NT>::kCtasPerSm)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,  // {3d, T, Bt} box {64,128,1}
                    const __grid_constant__ CUtensorMap tmO,    // {d, T, Bt}  box {64,32,1}
                    const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("gww: attention dynamic smem base not 1024-aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  using Cfg = AttnCfg<NT>;
  constexpr int kProducerWarp = 4 * NT, kMmaWarp = 4 * NT + 1;
  uint8_t* q_s = smem;                                  // NT x 16 KB (later: O staging)
  uint8_t* k_s = smem + Cfg::kQBytes;                   // kAttnStages x 16 KB
  uint8_t* v_s = k_s + kAttnStages * 16384;             // kAttnStages x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_s + kAttnStages * 16384);
  const uint32_t bar_q = smem_u32(bars);                // 1
  const uint32_t bar_kfull = bar_q + 8;                 // kAttnStages
  const uint32_t bar_kempty = bar_kfull + 8 * kAttnStages;
  const uint32_t bar_vfull = bar_kempty + 8 * kAttnStages;
  const uint32_t bar_vempty = bar_vfull + 8 * kAttnStages;
  const uint32_t bar_sfull = bar_vempty + 8 * kAttnStages;  // 2
  const uint32_t bar_pfull = bar_sfull + 16;                // 2
  const uint32_t bar_ofull = bar_pfull + 16;                // 2
  const uint32_t bar_sfree = bar_ofull + 16;                // 2 (softmax -> mma: S copied to registers)
  const uint32_t bar_pvdone = bar_sfree + 16;               // 2 (mma -> softmax: P.V finished)
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 1 + 4 * kAttnStages + 10);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int wg = warp >> 2;
  const int head = blockIdx.y, bi = blockIdx.z;
  const int q0 = blockIdx.x * (128 * NT);
  const int nkv = p.nkv;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(bar_q, 1);
    for (int i = 0; i < kAttnStages; ++i) {
      mbar_init(bar_kfull + 8 * i, 1);
      mbar_init(bar_kempty + 8 * i, 1);
      mbar_init(bar_vfull + 8 * i, 1);
      mbar_init(bar_vempty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_sfull + 8 * i, 1);
      mbar_init(bar_pfull + 8 * i, 4);       // one elected arrival per softmax warp
      mbar_init(bar_ofull + 8 * i, 1);
      mbar_init(bar_sfree + 8 * i, 4);
      mbar_init(bar_pvdone + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_ptr_s));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (wg == NT) {
    // register pool of the CTA (setmaxnreg draws from what the CTA's own warps release):
    //   NT = 2: 384 x 168 at launch; this group releases 128 x 72, the softmax groups take 256 x 32
    //   NT = 1: 256 x 128 at launch; this group releases 128 x 88, the softmax group takes 128 x 80
    if constexpr (NT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    // Both single-thread roles are written as warp-uniform loops in which one elected lane issues the
    // TMA / tcgen05 instructions (see elect_one() in ptx.cuh): the r1 profile showed the lone MMA
    // thread of the `lane == 0` version busy ~90 % of the time executing ~360 instructions per key
    // tile, i.e. Q.K^T and P.V were issued late and both softmax groups waited for them.
    if (warp == kProducerWarp) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, Cfg::kQBytes);
#pragma unroll
        for (int t = 0; t < NT; ++t) tma_load_3d(smem_u32(q_s + t * 16384), &tmQKV, bar_q, head * 64, q0 + t * 128, bi);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t k_smem = smem_u32(k_s), v_smem = smem_u32(v_s);
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
        if (++stage == kAttnStages) { stage = 0; phase ^= 1; }
      }
    } else if (warp == kMmaWarp) {
      // ===================== MMA issuer =====================
      constexpr uint32_t kIdescS = make_idesc_op16(128, 128, 0);
      constexpr uint32_t kIdescO = make_idesc_op16(128, 64, 1);   // V is MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      // TMEM map: S_t at t*128, P_t at NT*128 + t*64, O_t at NT*192 + t*64  (second entries unused for NT = 1)
      const uint32_t tS[2] = {tb + 0u, tb + 128u};
      const uint32_t tP[2] = {tb + NT * 128u, tb + NT * 128u + 64u};
      const uint32_t tO[2] = {tb + NT * 192u, tb + NT * 192u + 64u};
      const uint32_t q_smem = smem_u32(q_s), k_smem = smem_u32(k_s), v_smem = smem_u32(v_s);
      const uint64_t qdesc[2] = {make_sw128_desc(q_smem), make_sw128_desc(q_smem + 16384)};
      // S_t = Q_t K^T (4 MMAs of K=16) and its commit
      auto issue_s = [&](int t, const uint64_t kdesc) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tS[t], qdesc[t] + 2 * k, kdesc + 2 * k, kIdescS, k);
        umma_commit(bar_sfull + 8 * t);
      };
      // O_t (+)= P_t V (8 MMAs of K=16 keys) and its commit
      auto issue_pv = [&](int t, const uint64_t vdesc, bool first) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tO[t], tP[t] + 8 * k, vdesc + 128 * k, kIdescO, (!first || k) ? 1u : 0u);
        umma_commit(bar_pvdone + 8 * t);
      };
      mbar_wait(bar_q, 0);
      mbar_wait(bar_kfull, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t kdesc = make_sw128_desc(k_smem);
#pragma unroll
        for (int t = 0; t < NT; ++t) issue_s(t, kdesc);
        umma_commit(bar_kempty);
      }
      __syncwarp();
      // Event-driven issue: for each query tile t the warp polls (non-blocking) whether
      //   S_t(i)   can go: K tile i resident and softmax t has copied S_t(i-1) out of TMEM (s_free),
      //   P_t(i).V can go: V tile i resident and softmax t has written P_t(i) (p_full),
      // and issues whatever is ready.  The r1 fixed order (S0 | PV1 | S1 | PV0 with blocking waits)
      // coupled the two softmax groups: the leading group stalled on s_full ~500 clocks per tile
      // because its next Q.K^T sat behind the other group's p_full wait (ncu: 55 % of s_full waits retried).
      int s_next[2] = {1, 1}, pv_next[2] = {0, 0};
      int s_stage[2] = {1 % kAttnStages, 1 % kAttnStages}, pv_stage[2] = {0, 0};
      uint32_t s_phase[2] = {0, 0}, pv_phase[2] = {0, 0};
      if (kAttnStages == 1) { s_phase[0] = s_phase[1] = 1; }
      uint64_t idle_t0 = 0;
      while (pv_next[0] < nkv || (NT == 2 && pv_next[1] < nkv)) {
        bool progressed = false;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (s_next[t] < nkv) {
            const int i = s_next[t], st = s_stage[t];
            const uint32_t ok = mbar_test(bar_kfull + 8 * st, s_phase[t]) & mbar_test(bar_sfree + 8 * t, (i - 1) & 1);
            if (__all_sync(0xffffffffu, ok)) {
              tc_fence_after();
              if (elect_one()) {
                issue_s(t, make_sw128_desc(k_smem + st * 16384));
                if (NT == 1 || s_next[t ^ 1] > i) umma_commit(bar_kempty + 8 * st);   // all tiles' Q.K^T on K tile i issued
              }
              __syncwarp();
              s_next[t] = i + 1;
              if (++s_stage[t] == kAttnStages) { s_stage[t] = 0; s_phase[t] ^= 1; }
              progressed = true;
            }
          }
          if (pv_next[t] < nkv) {
            const int i = pv_next[t], st = pv_stage[t];
            const uint32_t ok = mbar_test(bar_vfull + 8 * st, pv_phase[t]) & mbar_test(bar_pfull + 8 * t, i & 1);
            if (__all_sync(0xffffffffu, ok)) {
              tc_fence_after();
              if (elect_one()) {
                issue_pv(t, make_sw128_desc(v_smem + st * 16384), i == 0);
                if (NT == 1 || pv_next[t ^ 1] > i) umma_commit(bar_vempty + 8 * st);  // all tiles' P.V on V tile i issued
                if (i + 1 == nkv) umma_commit(bar_ofull + 8 * t);
              }
              __syncwarp();
              pv_next[t] = i + 1;
              if (++pv_stage[t] == kAttnStages) { pv_stage[t] = 0; pv_phase[t] ^= 1; }
              progressed = true;
            }
          }
        }
        if (progressed) {
          idle_t0 = 0;
        } else {
          // nothing ready: back off briefly (none of these events is latency-critical: each is needed
          // about one key tile later) and bound the wait like mbar_wait does
          nanosleep_ns(64);
          const uint64_t now = globaltimer_ns();
          if (idle_t0 == 0) idle_t0 = now;
          else if (now - idle_t0 > GWW_MBAR_TIMEOUT_NS) {
            if (lane == 0)
              printf("gww: attention issue loop timeout block=(%d,%d,%d) s_next=(%d,%d) pv_next=(%d,%d)\n", blockIdx.x,
                     blockIdx.y, blockIdx.z, s_next[0], s_next[1], pv_next[0], pv_next[1]);
            __trap();
          }
        }
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    if constexpr (NT == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int t = wg;                           // query tile handled by this warpgroup
    const int wq = warp & 3;                    // TMEM lane quarter
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tP = tmem_base + lane_off + NT * 128 + t * 64;
    const uint32_t tO = tmem_base + lane_off + NT * 192 + t * 64;
    constexpr float kLog2e = 1.4426950408889634f;
    float m_used = 0.f, l = 0.f;
    const uint32_t b_sfull = bar_sfull + 8 * t, b_sfree = bar_sfree + 8 * t;
    const uint32_t b_pfull = bar_pfull + 8 * t, b_pvdone = bar_pvdone + 8 * t;

    // exp2 on the FMA/ALU pipes: n = rint(x) by the 1.5*2^23 trick, 2^f on [-0.5, 0.5] as a cubic,
    // exponent added to the bit pattern.
    auto exp2_fma = [](float x) {
      x = fmaxf(x, -126.0f);
      const float tt = x + 12582912.0f;
      const float f = x - (tt - 12582912.0f);
      float pl = fmaf(f, 0.05517164617776871f, 0.2426111251115799f);
      pl = fmaf(pl, f, 0.6932609677314758f);
      pl = fmaf(pl, f, 0.9999280571937561f);
      return __uint_as_float(__float_as_uint(pl) + (__float_as_uint(tt) << 23));
    };

    auto tile = [&](const int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      if (!(kAttnAblate & 16)) mbar_wait(b_sfull, j & 1);
      tc_fence_after();
      uint32_t s[4][32];
      if (!(kAttnAblate & 32)) {
        tmem_ld32(tS + 0, s[0]);
        tmem_ld32(tS + 32, s[1]);
        tmem_ld32(tS + 64, s[2]);
        tmem_ld32(tS + 96, s[3]);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c][i] = __float_as_uint(0.01f * (i + c + lane + j));
      }
      tc_fence_before();
      __syncwarp();                              // (32 per-thread arrivals on one barrier word serialise)
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
      const float mt = (kAttnAblate & 8) ? 4.0f : fmaxf(fmaxf(mt0, mt1), fmaxf(mt2, mt3));
      bool pv_waited = (j == 0);
      if (j == 0) {
        m_used = mt;
      } else {
        // lazy rescale: exact (same algebra as online softmax), but skipped while the stale max
        // keeps exp2 arguments <= 8.
        const bool need = (mt - m_used) * kLog2e > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(b_pvdone, (j - 1) & 1);      // O_t must be complete before it is rescaled
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
      const float mneg = -m_used * kLog2e;
      float l0 = 0.f, l1 = 0.f;
      uint32_t pk[32];
      constexpr int kGroups = (kAttnGroup > 0) ? 128 / kAttnGroup : 1;
      float lsnap[kGroups];                          // l0 after each group (compile-time indexed)
      float mneg_g = mneg;
      if (kAttnStagger && j == 0 && t == 1) named_bar_sync(2, 256);    // wait for group 0's half-tile mark
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (kAttnStagger && j == 0 && t == 0 && c == 2) named_bar_arrive(2, 256);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          if (kAttnGroup > 0) {
            const int e = c * 32 + i;                // element index within the tile
            if (e % kAttnGroup == 0) {
              const int g = e / kAttnGroup;
              if (g > 0) lsnap[g - 1] = l0;
              // == mneg (l0 is finite); a real data dependency for the scheduler
              mneg_g = (g >= kAttnLookahead) ? fmaf(lsnap[g - kAttnLookahead], 0.0f, mneg) : mneg;
            }
          }
          const float x0 = fmaf(__uint_as_float(s[c][i]), kLog2e, mneg_g);
          const float x1 = fmaf(__uint_as_float(s[c][i + 1]), kLog2e, mneg_g);
          // a fixed subset goes to the FMA pipes (never in the masked tile: -inf needs the MUFU)
          const bool f0 = !kMask && (i * kAttnFmaExp / 32) != ((i + 1) * kAttnFmaExp / 32);
          const bool f1 = !kMask && ((i + 1) * kAttnFmaExp / 32) != ((i + 2) * kAttnFmaExp / 32);
          const float p0 = (kAttnAblate & 2) ? x0 : (f0 ? exp2_fma(x0) : fast_exp2(x0));
          const float p1 = (kAttnAblate & 3) ? x1 : (f1 ? exp2_fma(x1) : fast_exp2(x1));
          if (!(kAttnAblate & 4) || i == 0) {
            l0 += p0;
            l1 += p1;
          }
          pk[(c & 1) * 16 + (i >> 1)] = pack_op16x2(p0, p1);
        }
        if (c & 1) {                                 // 64 columns packed -> 32 TMEM columns of P
          if (!pv_waited) {
            if (!(kAttnAblate & 16)) mbar_wait(b_pvdone, (j - 1) & 1);        // P_t(j-1).V finished reading P_t
            tc_fence_after();
            pv_waited = true;
          }
          if (!(kAttnAblate & 32)) tmem_st32(tP + (c >> 1) * 32, pk);
          else asm volatile("" ::"r"(pk[0] ^ pk[7] ^ pk[13] ^ pk[31]));
        }
      }
      l += l0 + l1;
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
    };
    for (int j = 0; j + 1 < nkv; ++j) tile(j, std::false_type{});
    if ((p.T & 127) != 0) tile(nkv - 1, std::true_type{});
    else tile(nkv - 1, std::false_type{});
    // ---- epilogue: O_t / l -> bf16 -> swizzled staging (the dead Q_t buffer) -> TMA store
    mbar_wait(bar_ofull + 8 * t, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    uint32_t o0[32], o1[32];
    tmem_ld32(tO, o0);
    tmem_ld32(tO + 32, o1);
    tmem_wait_ld();
    uint8_t* stg = q_s + t * 16384 + wq * 4096;
    uint8_t* sb = stg + lane * 128;
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
      tma_store_wait_read<0>();     // the CTA may exit once the staging buffer has been read
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace gww
