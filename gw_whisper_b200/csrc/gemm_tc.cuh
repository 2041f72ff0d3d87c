// Persistent, warp-specialised bf16 GEMM on tcgen05 / TMEM fed by TMA (sm_100a).
//
//   C[b, r, n] = epilogue( sum_{tap, c} A[b, r*? + tap, c] * W[n, tap*Cpad + c] )
//
// One kernel serves every dense contraction of the Whisper encoder
// (HF modeling_whisper.py:619-625 conv stem, :310/:331-332/:355 projections, :403-408 MLP):
//   * A is read through a 4-D tensor map {C, P, R, Bt}: for a plain Linear P=1 and taps=1; for the
//     k=3 convolutions the three taps are three row-shifted TMA boxes of the same (zero-row padded)
//     activation, i.e. im2col is done by the TMA unit, never materialised.
//   * W is [N, Ktot] bf16, K-major (nn.Linear layout), read through a 2-D map.
//   * Accumulators live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps
//     the MMAs of tile i+1.
//   * Epilogue warps read TMEM with tcgen05.ld (next 32-column chunk prefetched while the current one
//     is processed), apply bias / GELU / residual / positional embedding and write each thread's row
//     segment with 256-bit global stores (STG.256: one full 32-byte sector per instruction).  An
//     earlier version staged through smem + TMA stores; waiting for the TMA unit (shared with the
//     operand loads) to drain each 4 KB store made every f32-output GEMM epilogue-bound (r1 profile).
//
// Warp roles (384 threads): w0 = TMA producer, w1 = MMA issuer (one thread), w2 = TMEM allocator,
// w4..w11 = epilogue (TMEM lane quarter = warp_idx % 4, column half = (warp_idx - 4) / 4).  The
// encoder's GEMMs have K = 512..3072 against N up to 3072: per output element the tensor pipe needs
// only K/4096 clocks, so the epilogue (bias, GELU, residual traffic) must sustain ~5-8 outputs per
// clock per SM.  Two epilogue warps per SM sub-partition hide the TMEM / global-load latencies, and
// the f32 residual block of each tile is prefetched into L2 by the producer warp
// (cp.async.bulk.prefetch.tensor) one tile before the epilogue reads it.
#pragma once
#include "ptx.cuh"

namespace gww {

enum GemmEpilogue : int {
  EPI_BIAS_BF16 = 0,       // out bf16 = acc + bias                      (QKV projection)
  EPI_BIAS_GELU_BF16 = 1,  // out bf16 = gelu(acc + bias)                (conv1, fc1)
  EPI_BIAS_RESID_F32 = 2,  // out f32  = resid + acc + bias              (out_proj, fc2; in-place ok)
  EPI_BIAS_GELU_POS_F32 = 3  // out f32 = gelu(acc + bias) + pos[r, n]   (conv2 + embed_positions)
};

struct GemmParams {
  int rows;          // output rows per batch entry (R)
  int batch;         // Bt
  int n;             // output features (multiple of 64)
  int kb_per_tap;    // 64-wide K blocks per tap
  int taps;          // 1 (Linear) or 3 (conv k=3)
  int p_mod;         // P: tap -> (tap % P, tap / P) coordinates in dims 1 and 2 of the A map
  void* c;           // output base (bf16 or f32 depending on the epilogue)
  long c_row_stride;    // elements between consecutive rows
  long c_batch_stride;  // elements between batch entries
  const float* bias;   // [n] or nullptr
  const float* resid;  // [batch*rows, n] f32 (EPI_BIAS_RESID_F32)
  const float* pos;    // [rows, n] f32 (EPI_BIAS_GELU_POS_F32)
  int prefetch_resid;  // tmR is a valid {n, batch*rows} f32 map of resid: L2-prefetch it per tile
  // filled by the launcher: floor(2^40 / d) + 1 for d = tiles_n and d = tiles_per_batch (exact x / d for
  // x < 2^24, d < 2^16) -- the per-tile index math must not cost integer divisions in the single-thread loops
  unsigned long long magic_tiles_n;
  unsigned long long magic_tiles_per_batch;
  // ---- LayerNorm folding (see "LayerNorm folding" below).  Producer side (f32-output epilogues):
  op16_t* xb;     // bf16 copy of the output rows [batch*rows, n] (the next GEMM's A operand) or nullptr
  float* stats_out;      // per-row partial (sum, sum of squares): [batch*rows][stat_slots][2], slot = 2*n_tile + half
  // consumer side (bf16-output epilogues): out = rs * acc - rs * mu * c1[n] + c2[n], c2 passed as `bias`
  const float* stats_in; // the producer's partials for the rows of A, or nullptr (plain bias epilogue)
  const float* ln_c1;    // [n] column sums of the gamma-scaled (bf16-rounded) weights
  int stat_slots;        // partial slots per row written (producer) / to be summed (consumer)
  float ln_inv_k;        // 1 / d_model
  float ln_eps;
  float* ln_guard;       // device float (bits compared as uint): running max of |mean| / std over the rows consumed,
  float ln_guard_min;    //   recorded only above ln_guard_min (the fold loses precision when the residual stream has
                         //   a large common mode; the host switches the model to the stand-alone LayerNorm then)
};

// LayerNorm folding.  h = LN(x) feeds only a Linear: LN(x) W^T + b = rs * (x (g.W)^T) - rs * mu * c1 + c2 with
// c1[n] = sum_k g_k W_nk and c2[n] = sum_k beta_k W_nk + b_n.  So the GEMM that PRODUCES the residual stream x
// also writes its bf16 copy and per-row partial sums of x and x^2 (in fixed slots: deterministic, no atomics,
// no memset), and the consuming GEMM runs on the raw bf16 x with gamma-scaled weights and applies the two
// per-row scalars in its epilogue.  The stand-alone LayerNorm kernel (6 d bytes of HBM traffic per row, 7 % of
// the whisper-base step) disappears.  Rounding: bf16(x) against bf16(LN(x)) -- measured equal to within 5 %
// on every layer of the random-init and spread-scaled encoders (|mu| / sigma <= 1.2); models whose residual
// stream has |mu| >> sigma should keep the LayerNorm kernel (GWW_LN_FOLD=0).

__host__ __device__ __forceinline__ unsigned long long gemm_div_magic(int d) {
  return ((1ull << 40) / static_cast<unsigned long long>(d)) + 1ull;
}
__device__ __forceinline__ int gemm_fast_div(int x, unsigned long long magic) {
  return static_cast<int>((static_cast<unsigned long long>(x) * magic) >> 40);
}

// erf-GELU (HF ACT2FN["gelu"], modeling_whisper.py:403) in 8 instructions + 1 MUFU:
//   gelu(v) = 0.5 v (1 + erf(v/sqrt2)) ~= 0.5 v (1 + tanh(v (c1 + c2 v^2 + c3 v^4)))
// least-squares fit on [-6, 6]: |err| < 3.0e-5 absolute before the MUFU.TANH error (2^-11 relative),
// i.e. well below one bf16 ulp of the result (v^2 is clamped at 36 where tanh has saturated).
__device__ __forceinline__ float gelu_erf_fast(float v) {
  const float v2 = fminf(v * v, 36.0f);
  float p = fmaf(v2, -0.0003587323623918125f, 0.03705034510712041f);
  p = fmaf(v2, p, 0.7974584707758231f);
  const float th = fast_tanh(v * p);
  const float hv = 0.5f * v;
  return fmaf(hv, th, hv);
}

// The same GELU on a PAIR of values with packed fp32 math (FMUL2 / FFMA2): 7 packed + 2 FMNMX + 2 MUFU
// instructions per pair instead of 18; the GELU epilogues are bound by their issue slots.
//   in : packed pre-activation pair v;  out: gelu(v.lo), gelu(v.hi)
__device__ __forceinline__ void gelu_erf_fast2(const uint64_t v, float& g0, float& g1) {
  float q0, q1;
  f2_unpack(f2_mul(v, v), q0, q1);
  const uint64_t v2 = f2_pack(fminf(q0, 36.0f), fminf(q1, 36.0f));
  uint64_t p = f2_fma(v2, f2_pack(-0.0003587323623918125f, -0.0003587323623918125f),
                      f2_pack(0.03705034510712041f, 0.03705034510712041f));
  p = f2_fma(v2, p, f2_pack(0.7974584707758231f, 0.7974584707758231f));
  float t0, t1;
  f2_unpack(f2_mul(v, p), t0, t1);
  const uint64_t th = f2_pack(fast_tanh(t0), fast_tanh(t1));
  const uint64_t hv = f2_mul(v, f2_pack(0.5f, 0.5f));
  f2_unpack(f2_fma(hv, th, hv), g0, g1);
}

constexpr int kGemmThreads = 384;   // 12 warps: TMA, MMA, TMEM alloc, spare, 8 x epilogue
// The GELU -> bf16 epilogue (fc1, conv1) runs ~14 instructions per output pair.  GWW_GELU_EPI_WARPS=16 gives it
// 16 epilogue warps (12 for BN = 192; a quarter / third of the tile's columns each, single-buffered TMEM loads,
// <= 102 registers) instead of 8: measured equal in isolation (0.769 vs 0.772 ms) and slower inside the step
// (33.5 vs 30.9 ms), so 8 stays the default -- the kernel is not short of warps but of power budget.
#ifndef GWW_GELU_EPI_WARPS
#define GWW_GELU_EPI_WARPS 8
#endif
template <int EPI, int BN>
__host__ __device__ constexpr int gemm_epi_warps() {
  // column parts of whole 32-column chunks: 256 / 4, 192 / 3, 128 / 4
  return (EPI == EPI_BIAS_GELU_BF16 && GWW_GELU_EPI_WARPS == 16) ? (BN == 192 ? 12 : 16) : 8;
}
template <int EPI, int BN>
__host__ __device__ constexpr int gemm_threads() { return 128 + 32 * gemm_epi_warps<EPI, BN>(); }

template <int BN, int MC = 1>
struct GemmSmem {
  // MC == 2 (CTA pair): each CTA stages only its half of the weight tile
  static constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16
  static constexpr int kBBytes = (BN / MC) * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 192;  // (2*stages+4) mbarriers + tmem ptr, stages <= 6
  static constexpr int kStagingBytes = 8 * 4096;   // one 32-row x 128-byte transpose buffer per epilogue warp
  static constexpr int kBiasFloats = 3072;         // the whole bias vector lives in smem (N <= 3072)
  static constexpr int kBiasBytes = 2 * kBiasFloats * 4;   // bias (or c2) and the LayerNorm-fold column sums c1
  static constexpr int kLimit = 232448 - 1024;     // 227 KB opt-in minus the 1 KB reserved for __align__(1024)
  static constexpr int kFit = (kLimit - kStagingBytes - kBiasBytes - kBarBytes) / kStageBytes;
  static constexpr int kStages = kFit > 6 ? 6 : kFit;
  static_assert(kStages >= 3, "too few pipeline stages");
  // dynamic smem is declared __align__(1024); no slack needed (checked at kernel entry)
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBiasBytes + kBarBytes;
  static_assert(kTotal <= kLimit, "exceeds the dynamic shared memory limit of sm_100");
};

// MC = 2: a CTA pair (cluster of two, the two SMs of a TPC) computes a 256 x BN tile with ONE
// tcgen05.mma.cta_group::2 stream issued by the leader (cluster rank 0): CTA r owns rows [128r, 128r+128)
// of the tile (its A rows in its smem, its accumulator rows in its TMEM) and stages only weight rows
// [r*BN/2, (r+1)*BN/2) -- the tensor core reads the other half from the peer's shared memory.  With
// 128x256x64 single-CTA stages every SM ingests 48 KB per 512 tensor clocks (1 byte per 85 FLOP,
// ~12-13 TB/s of L2 -> SM traffic at 1 PFLOP/s: the measured crossbar ceiling, r1 ncu
// l1tex__m_xbar2l1tex_read_bytes), so the K >= 512 GEMMs were fabric-bound at ~60 % tensor-active; the
// pair needs 32 KB per SM for the same math.  (A TMA-multicast variant was measured first: no gain,
// since multicast saves L2 reads but not per-SM ingest.)
template <int BN, int EPI, int MC>
__global__ void __launch_bounds__(gemm_threads<EPI, BN>(), 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using S = GemmSmem<BN, MC>;
  constexpr int kStages = S::kStages;
  constexpr bool kOutF32 = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_GELU_POS_F32);
  constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_op16(128 * MC, BN, 0);
  constexpr uint16_t kPairMask = 0x3;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("gww: gemm dynamic smem base not 1024-aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  uint8_t* stage_base = smem;
  uint8_t* staging_base = smem + kStages * S::kStageBytes;
  float* bias_s = reinterpret_cast<float*>(staging_base + S::kStagingBytes);
  float* c1_s = bias_s + S::kBiasFloats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging_base + S::kStagingBytes + S::kBiasBytes);
  // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_tfull = bar_empty + 8 * kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;

  const int tiles_per_batch = (p.rows + 127) >> 7;
  const int tiles_m = tiles_per_batch * p.batch;
  const int tiles_n = (p.n + BN - 1) / BN;
  const int num_kb = p.kb_per_tap * p.taps;
  // work items: MC == 1 -> one tile per item; MC == 2 -> a vertical pair of tiles per cluster item
  const uint32_t cta_rank = (MC > 1) ? cluster_ctarank() : 0u;
  const int num_items = ((tiles_m + MC - 1) / MC) * tiles_n;
  const int item0 = blockIdx.x / MC;
  const int item_step = gridDim.x / MC;
  auto item_to_tile = [&](int item, int& m_idx, int& n_idx) {
    const int mp = gemm_fast_div(item, p.magic_tiles_n);
    n_idx = item - mp * tiles_n;
    m_idx = mp * MC + static_cast<int>(cta_rank);
    return m_idx < tiles_m;       // false: this CTA only takes part in the pair's loads and barriers
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  // the bias vector is read by every epilogue thread for every tile: keep it in shared memory (the r1
  // profile showed the per-chunk global bias loads missing L1 and stalling each chunk ~500 clocks)
  constexpr int kEpiWarps = gemm_epi_warps<EPI, BN>();
  for (int i = threadIdx.x; i < p.n; i += gemm_threads<EPI, BN>()) {
    bias_s[i] = (p.bias != nullptr) ? p.bias[i] : 0.0f;
    c1_s[i] = (p.stats_in != nullptr) ? p.ln_c1[i] : 0.0f;
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);       // pair: the leader's commit is multicast into both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, gemm_epi_warps<EPI, BN>() * MC);   // one elected arrival per epilogue warp (of both CTAs in pair mode)
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (MC == 1) {
      tmem_alloc<kTmemCols>(smem_u32(tmem_ptr_s));
      tmem_relinquish();
    } else {
      tmem_alloc_pair<kTmemCols>(smem_u32(tmem_ptr_s));   // same warp, same smem offset in both CTAs
      tmem_relinquish_pair();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // peer barriers exist before anything arrives at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_leader = (MC > 1) ? mapa_cluster(bar_full, 0) : bar_full;
      const uint32_t stage_smem = smem_u32(stage_base);
      for (int item = item0; item < num_items; item += item_step) {
        int m_idx, n_idx;
        const bool valid = item_to_tile(item, m_idx, n_idx);
        const int b = gemm_fast_div(m_idx, p.magic_tiles_per_batch);
        const int r0 = (m_idx - b * tiles_per_batch) << 7;
        if constexpr (EPI == EPI_BIAS_RESID_F32) {
          // pull this tile's residual block into L2 one tile ahead of the epilogue that adds it
          if (p.prefetch_resid && valid && elect_one()) tma_prefetch_l2_2d(&tmR, n_idx * BN, b * p.rows + r0);
          __syncwarp();
        }
        // (no integer divisions in this loop: a single thread issues it and must stay well ahead of
        //  the 512 tensor clocks one k-block lasts)
        const int n_row = n_idx * BN + ((MC > 1) ? static_cast<int>(cta_rank) * (BN / MC) : 0);
        int tap_p = 0, tap_r = r0, kcol = 0;       // (tap % P, r0 + tap / P), weight column of the k-block
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int kc = 0; kc < p.kb_per_tap; ++kc, kcol += 64) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t full = bar_full + 8 * stage;
            const uint32_t sa = stage_smem + stage * S::kStageBytes;
            if (elect_one()) {
              if constexpr (MC == 1) {
                mbar_arrive_expect_tx(full, S::kStageBytes);
                tma_load_4d(sa, &tmA, full, kc * 64, tap_p, tap_r, b);
                tma_load_2d(sa + S::kABytes, &tmB, full, kcol, n_row);
              } else {
                // both CTAs' boxes complete on the LEADER's full barrier, which expects the bytes of both
                if (cta_rank == 0) mbar_arrive_expect_tx(full, MC * S::kStageBytes);
                const uint32_t lfull = full_leader + 8 * stage;
                tma_load_4d_pair(sa, &tmA, lfull, kc * 64, tap_p, tap_r, b);   // OOB (invalid tile): zeros
                tma_load_2d_pair(sa + S::kABytes, &tmB, lfull, kcol, n_row);
              }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          if (++tap_p == p.p_mod) { tap_p = 0; ++tap_r; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair mode: the leader CTA only) =====================
    // warp-uniform loop; the elected lane issues the MMAs and commits
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint32_t stage_smem = smem_u32(stage_base);
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);   // warp-uniform for the compiler
      for (int item = item0; item < num_items; item += item_step) {
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base_u + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = stage_smem + stage * S::kStageBytes;
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sa + S::kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // +32 B per 16-element K step inside the 128B swizzle atom => +2 in the addr field
              if constexpr (MC == 1) umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
              else umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
            }
            if constexpr (MC == 1) umma_commit(bar_empty + 8 * stage);
            else umma_commit_pair(bar_empty + 8 * stage, kPairMask);   // frees the stage in both CTAs
            if (kb + 1 == num_kb) {
              if constexpr (MC == 1) umma_commit(bar_tfull + 8 * as);
              else umma_commit_pair(bar_tfull + 8 * as, kPairMask);    // accumulator halves ready in both CTAs
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps: lane quarter = warp % 4, column half = (warp-4) / 4)
    // tcgen05.ld hands every thread 32 consecutive columns of ITS row, so storing straight from the
    // registers makes each warp-wide access touch 32 different 128-byte lines: the r1 ncu capture
    // showed l1tex (tag lookups) as the busiest unit of every GEMM and the K=512 ones bound by it.
    // Each warp therefore transposes its 32x32 chunk through a private swizzled smem buffer and moves
    // whole 128-byte row segments per quarter-warp (4 lines per instruction instead of 32).
    const int e = warp - 4;
    const int q = e & 3;
    const int half = e >> 2;            // column part handled by this warp (0..kEpiWarps/4-1)
    constexpr int HW = BN / (kEpiWarps / 4);   // accumulator columns handled by one warp
    constexpr int NC = HW / 32;         // 32-column chunks per warp
    static_assert(HW % 32 == 0, "column part must be whole 32-column chunks");
    uint8_t* stg = staging_base + (e & 7) * 4096;   // (the 16-warp GELU epilogue does not stage)
    const uint32_t stg_row = smem_u32(stg) + lane * 128;          // this thread's row while staging
    const int srow = lane >> 3, sslot = lane & 7;                 // (row-in-pass, 16-byte slot) while moving
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t tempty_leader = (MC > 1) ? mapa_cluster(bar_tempty, 0) : bar_tempty;
    // all of this thread's tcgen05.ld of accumulator `a` have completed (tcgen05.wait::ld precedes)
    auto release_acc = [&](int a) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (MC == 1) mbar_arrive(bar_tempty + 8 * a);
        else mbar_arrive_cluster(tempty_leader + 8 * a);
      }
    };
    for (int item = item0; item < num_items; item += item_step) {
      int m_idx, n_idx;
      const bool valid = item_to_tile(item, m_idx, n_idx);
      const int b = valid ? gemm_fast_div(m_idx, p.magic_tiles_per_batch) : 0;
      const int r0 = valid ? (m_idx - b * tiles_per_batch) << 7 : 0;
      const int n0 = n_idx * BN + half * HW;
      const int rbase = r0 + q * 32;          // first row of this warp inside the batch entry
      const int rows_here = valid ? (p.rows - rbase) : 0;   // rows [0, rows_here) of this warp's 32 exist
      const float* bias_w = bias_s + n0;      // this warp's bias slice (may run past N for a ragged last tile)
      const float* c1_w = c1_s + n0;

      // LayerNorm fold: this thread's row scalars (rs, -rs*mu); (1, 0) makes the epilogue a plain bias add
      uint64_t rs2 = f2_pack(1.0f, 1.0f), nmr2 = f2_pack(0.0f, 0.0f);
      if constexpr (!kOutF32) {
        if (p.stats_in != nullptr) {
          float s1 = 0.f, s2 = 0.f;
          if (lane < rows_here) {
            const float2* st = reinterpret_cast<const float2*>(p.stats_in) +
                               (static_cast<size_t>(b) * p.rows + rbase + lane) * p.stat_slots;
            for (int i = 0; i < p.stat_slots; ++i) { const float2 v2 = __ldg(st + i); s1 += v2.x; s2 += v2.y; }
          }
          const float mu = s1 * p.ln_inv_k;
          const float var = fmaxf(s2 * p.ln_inv_k - mu * mu, 0.0f);
          const float rs = rsqrtf(var + p.ln_eps);
          rs2 = f2_pack(rs, rs);
          nmr2 = f2_pack(-rs * mu, -rs * mu);
          if (p.ln_guard != nullptr && n_idx == 0 && half == 0 && lane < rows_here) {
            const float ratio = fabsf(mu) * rs;
            if (ratio > p.ln_guard_min) atomicMax(reinterpret_cast<unsigned int*>(p.ln_guard), __float_as_uint(ratio));
          }
        }
      }
      mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * HW;
      uint32_t v[2][32];
      if constexpr (EPI == EPI_BIAS_GELU_BF16) {
        // GELU epilogues are issue-bound (measured: staging costs them 6 %): straight 256-bit stores
        const bool row_ok = lane < rows_here;
        op16_t* crow = reinterpret_cast<op16_t*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride +
                              static_cast<size_t>(row_ok ? rbase + lane : 0) * p.c_row_stride + n0;
        constexpr bool kDouble = (kEpiWarps == 8);   // 8 warps: next chunk's TMEM load in flight; 16: other warps hide it
        tmem_ld32(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int nc = n0 + c * 32;
          tmem_wait_ld();
          if (kDouble) {
            if (c + 1 < NC) tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
            else release_acc(as);
          }
          const uint32_t(&vc)[32] = v[kDouble ? (c & 1) : 0];
          uint32_t pk[16];
          if (nc < p.n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bq = *reinterpret_cast<const float4*>(bias_w + c * 32 + 4 * j);
              const float4 cq = *reinterpret_cast<const float4*>(c1_w + c * 32 + 4 * j);
              float a0, a1, a2, a3;
              gelu_erf_fast2(f2_fma(f2_pack(__uint_as_float(vc[4 * j]), __uint_as_float(vc[4 * j + 1])), rs2,
                                    f2_fma(f2_pack(cq.x, cq.y), nmr2, f2_pack(bq.x, bq.y))), a0, a1);
              gelu_erf_fast2(f2_fma(f2_pack(__uint_as_float(vc[4 * j + 2]), __uint_as_float(vc[4 * j + 3])), rs2,
                                    f2_fma(f2_pack(cq.z, cq.w), nmr2, f2_pack(bq.z, bq.w))), a2, a3);
              pk[2 * j] = pack_op16x2(a0, a1);
              pk[2 * j + 1] = pack_op16x2(a2, a3);
            }
          }
          if (!kDouble) {                            // the values are consumed: fetch the next chunk / free the accumulator
            if (c + 1 < NC) tmem_ld32(t_acc + (c + 1) * 32, v[0]);
            else release_acc(as);
          }
          if (nc < p.n && row_ok) {
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { lo[i] = pk[i]; hi[i] = pk[8 + i]; }
            st_global_v8(crow + c * 32, lo);
            st_global_v8(crow + c * 32 + 16, hi);
          }
        }
      } else if constexpr (!kOutF32) {
        op16_t* cb = reinterpret_cast<op16_t*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride +
                            static_cast<size_t>(rbase) * p.c_row_stride + n0;
        // move `ncols` (32 or 64) staged bf16 columns starting at column `col0` of the warp's range:
        // all shared loads first, then the stores (independent registers: the accesses overlap)
        auto flush = [&](int col0, int ncols) {
          __syncwarp();
          if (ncols == 64) {
            uint4 d[8];
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int row = 4 * ps + srow;
              d[ps] = *reinterpret_cast<const uint4*>(stg + row * 128 + ((sslot ^ (row & 7)) << 4));
            }
            op16_t* dst = cb + static_cast<size_t>(srow) * p.c_row_stride + col0 + 8 * sslot;
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              if (4 * ps + srow < rows_here) *reinterpret_cast<uint4*>(dst) = d[ps];
              dst += 4 * p.c_row_stride;
            }
          } else {
            const int sl = lane & 3, r8 = lane >> 2;
            uint4 d[4];
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const int row = 8 * ps + r8;
              d[ps] = *reinterpret_cast<const uint4*>(stg + row * 128 + ((sl ^ (row & 7)) << 4));
            }
            op16_t* dst = cb + static_cast<size_t>(r8) * p.c_row_stride + col0 + 8 * sl;
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              if (8 * ps + r8 < rows_here) *reinterpret_cast<uint4*>(dst) = d[ps];
              dst += 8 * p.c_row_stride;
            }
          }
          __syncwarp();
        };
        tmem_ld32(t_acc, v[0]);
        int pending = 0;                      // staged, not yet flushed columns
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int nc = n0 + c * 32;
          float4 bv[8], cv[8];
          if (nc < p.n) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              bv[j] = *reinterpret_cast<const float4*>(bias_w + c * 32 + 4 * j);
              cv[j] = *reinterpret_cast<const float4*>(c1_w + c * 32 + 4 * j);
            }
          }
          tmem_wait_ld();
          if (c + 1 < NC) {
            tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
          } else {
            release_acc(as);                       // accumulator drained -> MMA may reuse it
          }
          if (nc < p.n) {                          // (N is a multiple of 64: chunk fully in or out)
            const uint32_t(&vc)[32] = v[c & 1];
            const int hslot = (pending == 32) ? 4 : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 bq = bv[2 * j + h], cq = cv[2 * j + h];
                float a0, a1, a2, a3;     // rs * acc + (-rs*mu * c1 + c2); (rs, -rs*mu) = (1, 0) without the fold
                f2_unpack(f2_fma(f2_pack(__uint_as_float(vc[8 * j + 4 * h]), __uint_as_float(vc[8 * j + 4 * h + 1])), rs2,
                                 f2_fma(f2_pack(cq.x, cq.y), nmr2, f2_pack(bq.x, bq.y))), a0, a1);
                f2_unpack(f2_fma(f2_pack(__uint_as_float(vc[8 * j + 4 * h + 2]), __uint_as_float(vc[8 * j + 4 * h + 3])), rs2,
                                 f2_fma(f2_pack(cq.z, cq.w), nmr2, f2_pack(bq.z, bq.w))), a2, a3);
                pk[2 * h] = pack_op16x2(a0, a1);
                pk[2 * h + 1] = pack_op16x2(a2, a3);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stg_row + (((hslot + j) ^ (lane & 7)) << 4)),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
            }
            pending += 32;
            if (pending == 64) { flush(c * 32 - 32, 64); pending = 0; }
          }
          if (c + 1 == NC && pending == 32) { flush(c * 32 - (nc < p.n ? 0 : 32), 32); pending = 0; }
        }
      } else {
        // residual / positional addend: fetched in the MOVE layout (coalesced), one chunk ahead; the
        // residual epilogue folds the bias into it (each thread owns 4 fixed columns per chunk)
        const float* addb;
        if constexpr (EPI == EPI_BIAS_RESID_F32) addb = p.resid + (static_cast<size_t>(b) * p.rows + rbase + srow) * p.n + n0 + 4 * sslot;
        else addb = p.pos + static_cast<size_t>(rbase + srow) * p.n + n0 + 4 * sslot;
        float* cb = reinterpret_cast<float*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride +
                    static_cast<size_t>(rbase + srow) * p.c_row_stride + n0 + 4 * sslot;
        float4 add[2][8];
        auto load_add = [&](int c, float4 (&dst)[8]) {
          const float* src = addb + c * 32;
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            dst[ps] = (4 * ps + srow < rows_here) ? *reinterpret_cast<const float4*>(src) : make_float4(0.f, 0.f, 0.f, 0.f);
            src += 4 * static_cast<size_t>(p.n);
          }
          if constexpr (EPI == EPI_BIAS_RESID_F32) {
            const float4 bq = *reinterpret_cast<const float4*>(bias_w + c * 32 + 4 * sslot);
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) { dst[ps].x += bq.x; dst[ps].y += bq.y; dst[ps].z += bq.z; dst[ps].w += bq.w; }
          }
        };
        op16_t* xbp = (p.xb != nullptr)
                                 ? p.xb + (static_cast<size_t>(b) * p.rows + rbase + srow) * p.n + n0 + 4 * sslot
                                 : nullptr;
        float rsum[8], rsq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { rsum[i] = 0.f; rsq[i] = 0.f; }
        tmem_ld32(t_acc, v[0]);
        load_add(0, add[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float4 bv[8];
          if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bv[j] = *reinterpret_cast<const float4*>(bias_w + c * 32 + 4 * j);
          }
          tmem_wait_ld();
          if (c + 1 < NC) {
            tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
            load_add(c + 1, add[(c + 1) & 1]);
          } else {
            release_acc(as);
          }
          const uint32_t(&vc)[32] = v[c & 1];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a0 = __uint_as_float(vc[4 * j]), a1 = __uint_as_float(vc[4 * j + 1]);
            float a2 = __uint_as_float(vc[4 * j + 2]), a3 = __uint_as_float(vc[4 * j + 3]);
            if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
              gelu_erf_fast2(f2_add(f2_pack(a0, a1), f2_pack(bv[j].x, bv[j].y)), a0, a1);
              gelu_erf_fast2(f2_add(f2_pack(a2, a3), f2_pack(bv[j].z, bv[j].w)), a2, a3);
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg_row + ((j ^ (lane & 7)) << 4)),
                         "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          }
          __syncwarp();
          float* dst = cb + c * 32;
#pragma unroll
          for (int g4 = 0; g4 < 2; ++g4) {       // 4 shared loads in flight, then their 4 stores
            float4 d[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int row = 4 * (4 * g4 + k) + srow;
              d[k] = *reinterpret_cast<const float4*>(stg + row * 128 + ((sslot ^ (row & 7)) << 4));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 ad = add[c & 1][4 * g4 + k];
              d[k].x += ad.x; d[k].y += ad.y; d[k].z += ad.z; d[k].w += ad.w;
              const bool rok = 4 * (4 * g4 + k) + srow < rows_here;
              if (rok) *reinterpret_cast<float4*>(dst) = d[k];
              dst += 4 * p.c_row_stride;
              if (xbp != nullptr) {
                if (rok) *reinterpret_cast<uint2*>(xbp + static_cast<size_t>(4 * (4 * g4 + k)) * p.n + c * 32) =
                    make_uint2(pack_op16x2(d[k].x, d[k].y), pack_op16x2(d[k].z, d[k].w));
                rsum[4 * g4 + k] += (d[k].x + d[k].y) + (d[k].z + d[k].w);
                rsq[4 * g4 + k] += fmaf(d[k].x, d[k].x, d[k].y * d[k].y) + fmaf(d[k].z, d[k].z, d[k].w * d[k].w);
              }
            }
          }
          __syncwarp();
        }
        if (p.stats_out != nullptr) {
          // this warp's partial over its HW columns: combine the 8 lanes that share a row, store to the slot
#pragma unroll
          for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              rsum[i] += __shfl_xor_sync(0xffffffffu, rsum[i], o);
              rsq[i] += __shfl_xor_sync(0xffffffffu, rsq[i], o);
            }
          }
          if (sslot == 0) {
            const int slot = 2 * n_idx + half;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = 4 * i + srow;
              if (row < rows_here)
                reinterpret_cast<float2*>(p.stats_out)[(static_cast<size_t>(b) * p.rows + rbase + row) * p.stat_slots + slot] =
                    make_float2(rsum[i], rsq[i]);
            }
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // no CTA leaves while the pair's MMAs / arrivals may still touch it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (MC == 1) tmem_dealloc<kTmemCols>(tmem_base);
    else tmem_dealloc_pair<kTmemCols>(tmem_base);
  }
}

}  // namespace gww
