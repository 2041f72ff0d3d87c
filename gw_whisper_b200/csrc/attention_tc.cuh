// Fused non-causal multi-head self-attention for the Whisper encoder on tcgen05 / TMEM (sm_100a).
// Reference math: HF modeling_whisper.py:215-238 (eager_attention_forward, scaling=1.0, no mask,
// no dropout) with q pre-scaled by head_dim^-0.5 (:310) -- the scale is folded into W_q / b_q at
// model-load time, so this kernel computes softmax(Q K^T) V per (det-window, head).
//
// Layout: qkv [Bt, T, 3*d] bf16 (q | k | v, head h at columns h*64), out [Bt, T, d] bf16.
// One CTA = one (det-window, head, pair of 128-row query tiles); it streams the T keys in 128-row
// K/V tiles through a 3-stage TMA ring.
//   S_t = Q_t K_j^T     tcgen05.mma SS, M=128 N=128 K=64, accumulator in TMEM (fp32)
//   softmax             FOUR warpgroups: query tile t in {0,1} x key half h in {0,1}; a thread owns one
//                       row and 64 of the tile's 128 keys; the row maximum is combined across the two
//                       halves through shared memory + a 256-thread named barrier; online max with
//                       lazy rescale (only when the running max grows by > 2^8), exp2 on MUFU
//   P_t -> TMEM (bf16)  tcgen05.st, then O_t += P_t V_j as tcgen05.mma TS (A from TMEM, V is the
//                       MN-major B operand straight from its TMA tile), M=128 N=64 K=128
// TMEM map (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O0 [384,448) O1 [448,512)
// Warps (640 threads): WG0..WG3 = softmax (t = wg >> 1, h = wg & 1), WG4 = {w16: TMA producer,
// w17: MMA issuer + TMEM alloc}.
//
// Scheduling (r1 measurements: MUFU.EX2 is the binding pipe at 16 results/clk/SM):
//   * v3 had one thread per row over all 128 keys, i.e. two softmax warps per SM sub-partition.
//     ptxas emits each warp's exponentials as one burst of back-to-back MUFUs, so with two warps the
//     pipe idled whenever both were in their load / max / FADD / store phases (62 % MUFU-active).
//     Four warps per sub-partition, each with half the keys (and ~100 instead of ~170 registers),
//     let the hardware interleave one warp's MUFU burst with the others' FMA-pipe work;
//   * S_t(j+1) is issued as soon as both halves have copied S_t(j) into registers (s_free): the
//     tensor work runs a whole tile ahead of the exponentials and is completely hidden;
//   * lazy rescale: O and l are rescaled only when the tile max exceeds the running max by more
//     than 2^8 (rare); exact online-softmax algebra either way; each half rescales its 32 O columns;
//   * P_t(j).V(j) is issued when both halves of P are written; softmax only waits for its completion
//     (pv_done) right before it overwrites P / rescales O for tile j+1; key masking exists only in the
//     last tile's code.
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
constexpr int kAttnThreads = 640;
// Q (later O staging) | K ring | V ring | row-max exchange [2 parities][2 tiles][2 halves][128] f32 |
// row-sum exchange [2 tiles][2 halves][128] f32 | barriers
constexpr int kAttnSmemBytes = 32768 + kAttnStages * 2 * 16384 + 4096 + 2048 + 256;

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV,  // {3d, T, Bt} box {64,128,1}
                    const __grid_constant__ CUtensorMap tmO,    // {d, T, Bt}  box {64,32,1}
                    const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("gww: attention dynamic smem base not 1024-aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  uint8_t* q_s = smem;                                  // 2 x 16 KB (later: O staging)
  uint8_t* k_s = smem + 32768;                          // kAttnStages x 16 KB
  uint8_t* v_s = k_s + kAttnStages * 16384;             // kAttnStages x 16 KB
  float* mx_s = reinterpret_cast<float*>(v_s + kAttnStages * 16384);   // [2][2][2][128]
  float* ls_s = mx_s + 1024;                                            // [2][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ls_s + 512);
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
  const int qpair = blockIdx.x, head = blockIdx.y, bi = blockIdx.z;
  const int q0 = qpair * 256;
  const int nkv = p.nkv;

  if (warp == 16 && lane == 0) {
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
      mbar_init(bar_pfull + 8 * i, 256);     // both key halves of the tile
      mbar_init(bar_ofull + 8 * i, 1);
      mbar_init(bar_sfree + 8 * i, 256);
      mbar_init(bar_pvdone + 8 * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 17) {
    tmem_alloc<512>(smem_u32(tmem_ptr_s));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (wg == 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    // Both single-thread roles are written as warp-uniform loops in which one elected lane issues the
    // TMA / tcgen05 instructions (see elect_one() in ptx.cuh): the r1 profile showed the lone MMA
    // thread of the `lane == 0` version busy ~90 % of the time executing ~360 instructions per key
    // tile, i.e. Q.K^T and P.V were issued late and both softmax groups waited for them.
    if (warp == 16) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_q, 32768);
        tma_load_3d(smem_u32(q_s), &tmQKV, bar_q, head * 64, q0, bi);
        tma_load_3d(smem_u32(q_s + 16384), &tmQKV, bar_q, head * 64, q0 + 128, bi);
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
    } else if (warp == 17) {
      // ===================== MMA issuer =====================
      constexpr uint32_t kIdescS = make_idesc_bf16(128, 128, 0);
      constexpr uint32_t kIdescO = make_idesc_bf16(128, 64, 1);   // V is MN-major
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      const uint32_t tS[2] = {tb + 0u, tb + 128u};
      const uint32_t tP[2] = {tb + 256u, tb + 320u};
      const uint32_t tO[2] = {tb + 384u, tb + 448u};
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
        issue_s(0, kdesc);
        issue_s(1, kdesc);
        umma_commit(bar_kempty);
      }
      __syncwarp();
      // issue order per step j:  S0(j+1) | P1(j-1).V | S1(j+1) | P0(j).V   (matches the order in
      // which the softmax groups produce their events; any other order is still safe)
      int stage = 0;              // stage of K/V tile j
      uint32_t phase = 0;
      int pstage = 0;             // stage of V tile j-1
      for (int j = 0; j < nkv; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == kAttnStages) { nstage = 0; nphase ^= 1; }
        const bool has_next = (j + 1 < nkv);
        const uint64_t kdesc_n = make_sw128_desc(k_smem + nstage * 16384);
        if (has_next) {
          mbar_wait(bar_kfull + 8 * nstage, nphase);
          mbar_wait(bar_sfree + 8 * 0, j & 1);
          tc_fence_after();
          if (elect_one()) issue_s(0, kdesc_n);
          __syncwarp();
        }
        if (j >= 1) {
          mbar_wait(bar_pfull + 8 * 1, (j - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(1, make_sw128_desc(v_smem + pstage * 16384), j - 1 == 0);
            umma_commit(bar_vempty + 8 * pstage);
          }
          __syncwarp();
        }
        if (has_next) {
          mbar_wait(bar_sfree + 8 * 1, j & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_s(1, kdesc_n);
            umma_commit(bar_kempty + 8 * nstage);
          }
          __syncwarp();
        }
        mbar_wait(bar_vfull + 8 * stage, phase);
        mbar_wait(bar_pfull + 8 * 0, j & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(0, make_sw128_desc(v_smem + stage * 16384), j == 0);
          if (!has_next) umma_commit(bar_ofull + 8 * 0);
        }
        __syncwarp();
        pstage = stage;
        stage = nstage;
        phase = nphase;
      }
      mbar_wait(bar_pfull + 8 * 1, (nkv - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(1, make_sw128_desc(v_smem + pstage * 16384), nkv - 1 == 0);
        umma_commit(bar_vempty + 8 * pstage);
        umma_commit(bar_ofull + 8 * 1);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax warpgroups =====================
    // register pool of the CTA: 640 x 96 at launch; WG4 releases 128 x (96 - 32) = 8192, the four
    // softmax warpgroups take 512 x (112 - 96) = 8192 (all) of them (setmaxnreg draws from the CTA's own pool)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int t = wg >> 1;                      // query tile handled by this warpgroup
    const int h = wg & 1;                       // key half of every 128-key tile
    const int wq = warp & 3;                    // TMEM lane quarter
    const int row = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128 + h * 64;
    const uint32_t tP = tmem_base + lane_off + 256 + t * 64 + h * 32;
    const uint32_t tO = tmem_base + lane_off + 384 + t * 64 + h * 32;   // this thread's 32 of the 64 O columns
    constexpr float kLog2e = 1.4426950408889634f;
    float m_used = 0.f, l = 0.f;
    const uint32_t b_sfull = bar_sfull + 8 * t, b_sfree = bar_sfree + 8 * t;
    const uint32_t b_pfull = bar_pfull + 8 * t, b_pvdone = bar_pvdone + 8 * t;
    float* mx_mine = mx_s + (t * 2 + h) * 128 + row;
    const float* mx_other = mx_s + (t * 2 + (h ^ 1)) * 128 + row;
    const uint32_t pair_bar = 1 + t;            // named barrier of the two warpgroups of tile t

    auto tile = [&](const int j, auto mask_tag) {
      constexpr bool kMask = decltype(mask_tag)::value;
      mbar_wait(b_sfull, j & 1);
      tc_fence_after();
      uint32_t s[2][32];
      tmem_ld32(tS + 0, s[0]);
      tmem_ld32(tS + 32, s[1]);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(b_sfree);                      // this half of S_t is in registers
      if constexpr (kMask) {
        const int valid = p.T - j * 128 - h * 64;   // keys [0, valid) of this half exist
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) s[c][i] = 0xff800000u;  // -inf
      }
      float mt0 = -INFINITY, mt1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mt0 = fmaxf(mt0, __uint_as_float(s[0][i]));
        mt1 = fmaxf(mt1, __uint_as_float(s[1][i]));
      }
      const float mp = fmaxf(mt0, mt1);
      mx_mine[(j & 1) * 512] = mp;               // double-buffered by tile parity
      named_bar_sync(pair_bar, 256);
      const float mt = fmaxf(mp, mx_other[(j & 1) * 512]);
      bool pv_waited = (j == 0);
      if (j == 0) {
        m_used = mt;
      } else {
        // lazy rescale: exact (same algebra as online softmax), but skipped while the stale max
        // keeps exp2 arguments <= 8.  Both halves of a row take the same decision (same mt, m_used).
        const bool need = (mt - m_used) * kLog2e > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          mbar_wait(b_pvdone, (j - 1) & 1);      // O_t must be complete before it is rescaled
          tc_fence_after();
          pv_waited = true;
          const float sc = need ? fast_exp2((m_used - mt) * kLog2e) : 1.0f;
          uint32_t o[32];
          tmem_ld32(tO, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
          tmem_st32(tO, o);
          tmem_wait_st();
          l *= sc;
          if (need) m_used = mt;
        }
      }
      const float mneg = -m_used * kLog2e;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[c][i]), kLog2e, mneg));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[c][i + 1]), kLog2e, mneg));
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        if (!pv_waited) {
          mbar_wait(b_pvdone, (j - 1) & 1);      // P_t(j-1).V finished reading P_t
          tc_fence_after();
          pv_waited = true;
        }
        tmem_st16(tP + c * 16, pk);              // 32 keys -> 16 TMEM columns of bf16 pairs
      }
      l += l0 + l1;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(b_pfull);
    };
    for (int j = 0; j + 1 < nkv; ++j) tile(j, std::false_type{});
    if ((p.T & 127) != 0) tile(nkv - 1, std::true_type{});
    else tile(nkv - 1, std::false_type{});
    // ---- epilogue: O_t / l -> bf16 -> swizzled staging (the dead Q_t buffer) -> TMA store
    ls_s[(t * 2 + h) * 128 + row] = l;
    named_bar_sync(pair_bar, 256);
    const float inv_l = 1.0f / (l + ls_s[(t * 2 + (h ^ 1)) * 128 + row]);
    mbar_wait(bar_ofull + 8 * t, 0);
    tc_fence_after();
    uint32_t o[32];
    tmem_ld32(tO, o);
    tmem_wait_ld();
    uint8_t* stg = q_s + t * 16384 + wq * 4096;
    uint8_t* sb = stg + lane * 128;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        pk[k] = pack_bf16x2(__uint_as_float(o[e * 8 + 2 * k]) * inv_l, __uint_as_float(o[e * 8 + 2 * k + 1]) * inv_l);
      const int jc = h * 4 + e;                  // 16-byte chunk of the 128-byte output row
      *reinterpret_cast<uint4*>(sb + ((jc ^ (lane & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    fence_proxy_async_smem();
    named_bar_sync(pair_bar, 256);               // both column halves of the 32-row boxes are staged
    if (h == 0 && lane == 0) {
      tma_store_3d(&tmO, smem_u32(stg), head * 64, q0 + t * 128 + wq * 32, bi);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace gww
