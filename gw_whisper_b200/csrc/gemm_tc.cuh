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
};

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

constexpr int kGemmThreads = 384;   // 12 warps: TMA, MMA, TMEM alloc, spare, 8 x epilogue

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : ((BN == 192) ? 4 : 6);
  static constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 192;  // (2*stages+4) mbarriers + tmem ptr, stages <= 6
  static constexpr int kStagingBytes = 8 * 4096;   // one 32-row x 128-byte transpose buffer per epilogue warp
  // dynamic smem is declared __align__(1024); no slack needed (checked at kernel entry)
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarBytes;
  // 227 KB opt-in limit minus the 1 KB the compiler reserves statically for the __align__(1024)
  static_assert(kTotal <= 232448 - 1024, "exceeds the dynamic shared memory limit of sm_100");
};

// MC = 2: clusters of two CTAs work on vertically adjacent tiles (same n, m and m+1) in lockstep and
// share the weight tile: each CTA fetches half of it and TMA-multicasts it into both, cutting the
// L2 -> SM traffic per k-block from 48 KB to 32 KB (BN = 256).  With 128x256x64 stages the kernel moves
// 1 byte per 85 FLOP, i.e. ~11.7 TB/s at 1 PFLOP/s -- the measured L2 ceiling (r1 ncu) -- so the
// big-K GEMMs were L2-bound, not tensor-bound.
template <int BN, int EPI, int MC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr bool kOutF32 = (EPI == EPI_BIAS_RESID_F32 || EPI == EPI_BIAS_GELU_POS_F32);
  constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t kIdesc = make_idesc_bf16(128, BN, 0);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) printf("gww: gemm dynamic smem base not 1024-aligned (0x%x)\n", smem_u32(smem));
    __trap();
  }
  uint8_t* stage_base = smem;
  uint8_t* staging_base = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging_base + S::kStagingBytes);
  // barrier layout: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then tmem ptr
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * kStages;
  const uint32_t bar_tfull = bar_empty + 8 * kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
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
    const int mp = item / tiles_n;
    n_idx = item - mp * tiles_n;
    m_idx = mp * MC + static_cast<int>(cta_rank);
    return m_idx < tiles_m;       // false: this CTA only takes part in the pair's loads and barriers
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, MC);      // every CTA of the cluster releases a stage into all of them
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<kTmemCols>(smem_u32(tmem_ptr_s));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // peer barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        int m_idx, n_idx;
        const bool valid = item_to_tile(item, m_idx, n_idx);
        const int b = m_idx / tiles_per_batch;
        const int r0 = (m_idx - b * tiles_per_batch) << 7;
        if constexpr (EPI == EPI_BIAS_RESID_F32) {
          // pull this tile's residual block into L2 one tile ahead of the epilogue that adds it
          if (p.prefetch_resid && valid) tma_prefetch_l2_2d(&tmR, n_idx * BN, b * p.rows + r0);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, S::kStageBytes);
          const uint32_t sa = smem_u32(stage_base + stage * S::kStageBytes);
          const int tap = kb / p.kb_per_tap;
          const int kc = kb - tap * p.kb_per_tap;
          tma_load_4d(sa, &tmA, full, kc * 64, tap % p.p_mod, r0 + tap / p.p_mod, b);   // OOB (invalid tile): zeros
          if constexpr (MC == 1) {
            tma_load_2d(sa + S::kABytes, &tmB, full, kb * 64, n_idx * BN);
          } else {
            constexpr int HB = BN / MC;         // weight rows fetched by this CTA, multicast to the pair
            tma_load_2d_mc(sa + S::kABytes + cta_rank * (HB * 128), &tmB, full, kb * 64,
                           n_idx * BN + static_cast<int>(cta_rank) * HB, static_cast<uint16_t>((1u << MC) - 1u));
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * S::kStageBytes);
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 B per 16-element K step inside the 128B swizzle atom => +2 in the addr field
            umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
          }
          if constexpr (MC == 1) umma_commit(bar_empty + 8 * stage);
          else umma_commit_mc(bar_empty + 8 * stage, static_cast<uint16_t>((1u << MC) - 1u));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(bar_tfull + 8 * as);
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
    const int half = e >> 2;
    constexpr int HW = BN / 2;          // accumulator columns handled by one warp
    constexpr int NC = HW / 32;         // 32-column chunks per warp
    uint8_t* stg = staging_base + e * 4096;
    const uint32_t stg_row = smem_u32(stg) + lane * 128;          // this thread's row while staging
    const int srow = lane >> 3, sslot = lane & 7;                 // (row-in-pass, 16-byte slot) while moving
    int as = 0;
    uint32_t aphase = 0;
    for (int item = item0; item < num_items; item += item_step) {
      int m_idx, n_idx;
      const bool valid = item_to_tile(item, m_idx, n_idx);
      const int b = valid ? m_idx / tiles_per_batch : 0;
      const int r0 = valid ? (m_idx - b * tiles_per_batch) << 7 : 0;
      const int n0 = n_idx * BN + half * HW;
      const int rbase = r0 + q * 32;          // first row of this warp inside the batch entry

      mbar_wait(bar_tfull + 8 * as, aphase);
      tc_fence_after();
      const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half * HW;
      uint32_t v[2][32];
      if constexpr (EPI == EPI_BIAS_GELU_BF16) {
        // GELU epilogues are issue-bound (measured: staging costs them 6 %): straight 256-bit stores
        const int r = rbase + lane;
        const bool row_ok = valid && r < p.rows;
        __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride +
                              static_cast<size_t>(row_ok ? r : 0) * p.c_row_stride + n0;
        tmem_ld32(t_acc, v[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          tmem_wait_ld();
          if (c + 1 < NC) {
            tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
          } else {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * as);
          }
          const int nc = n0 + c * 32;
          if (nc < p.n) {
            const uint32_t(&vc)[32] = v[c & 1];
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc);
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bv = (p.bias != nullptr) ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float a0 = gelu_erf_fast(__uint_as_float(vc[4 * j]) + bv.x);
              const float a1 = gelu_erf_fast(__uint_as_float(vc[4 * j + 1]) + bv.y);
              const float a2 = gelu_erf_fast(__uint_as_float(vc[4 * j + 2]) + bv.z);
              const float a3 = gelu_erf_fast(__uint_as_float(vc[4 * j + 3]) + bv.w);
              pk[2 * j] = pack_bf16x2(a0, a1);
              pk[2 * j + 1] = pack_bf16x2(a2, a3);
            }
            if (row_ok) {
              uint32_t lo[8], hi[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) { lo[i] = pk[i]; hi[i] = pk[8 + i]; }
              st_global_v8(crow + c * 32, lo);
              st_global_v8(crow + c * 32 + 16, hi);
            }
          }
        }
      } else if constexpr (!kOutF32) {
        __nv_bfloat16* cb = reinterpret_cast<__nv_bfloat16*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride + n0;
        // move `ncols` (32 or 64) staged bf16 columns starting at column `col0` of the warp's range
        auto flush = [&](int col0, int ncols) {
          __syncwarp();
          if (ncols == 64) {
#pragma unroll
            for (int ps = 0; ps < 8; ++ps) {
              const int row = 4 * ps + srow;
              const uint4 d = *reinterpret_cast<const uint4*>(stg + row * 128 + ((sslot ^ (row & 7)) << 4));
              const int rr = rbase + row;
              if (valid && rr < p.rows)
                *reinterpret_cast<uint4*>(cb + static_cast<size_t>(rr) * p.c_row_stride + col0 + 8 * sslot) = d;
            }
          } else {
#pragma unroll
            for (int ps = 0; ps < 4; ++ps) {
              const int row = 8 * ps + (lane >> 2), sl = lane & 3;
              const uint4 d = *reinterpret_cast<const uint4*>(stg + row * 128 + ((sl ^ (row & 7)) << 4));
              const int rr = rbase + row;
              if (valid && rr < p.rows)
                *reinterpret_cast<uint4*>(cb + static_cast<size_t>(rr) * p.c_row_stride + col0 + 8 * sl) = d;
            }
          }
          __syncwarp();
        };
        tmem_ld32(t_acc, v[0]);
        int pending = 0;                      // staged, not yet flushed columns
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          tmem_wait_ld();
          if (c + 1 < NC) {
            tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
          } else {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * as);      // accumulator drained -> MMA may reuse it
          }
          const int nc = n0 + c * 32;
          if (nc < p.n) {                          // (N is a multiple of 64: chunk fully in or out)
            const uint32_t(&vc)[32] = v[c & 1];
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + nc);
            const int hslot = (pending == 32) ? 4 : 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t pk[4];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 bv = (p.bias != nullptr) ? __ldg(b4 + 2 * j + h) : make_float4(0.f, 0.f, 0.f, 0.f);
                float a0 = __uint_as_float(vc[8 * j + 4 * h]) + bv.x, a1 = __uint_as_float(vc[8 * j + 4 * h + 1]) + bv.y;
                float a2 = __uint_as_float(vc[8 * j + 4 * h + 2]) + bv.z, a3 = __uint_as_float(vc[8 * j + 4 * h + 3]) + bv.w;
                if constexpr (EPI == EPI_BIAS_GELU_BF16) {
                  a0 = gelu_erf_fast(a0); a1 = gelu_erf_fast(a1);
                  a2 = gelu_erf_fast(a2); a3 = gelu_erf_fast(a3);
                }
                pk[2 * h] = pack_bf16x2(a0, a1);
                pk[2 * h + 1] = pack_bf16x2(a2, a3);
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
        // residual / positional addend: fetched in the MOVE layout (coalesced), one chunk ahead
        const float* addb;
        if constexpr (EPI == EPI_BIAS_RESID_F32) addb = p.resid + static_cast<size_t>(b) * p.rows * p.n + n0;
        else addb = p.pos + n0;
        float* cb = reinterpret_cast<float*>(p.c) + static_cast<size_t>(b) * p.c_batch_stride + n0;
        float4 add[2][8];
        auto load_add = [&](int c, float4 (&dst)[8]) {
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int rr = rbase + 4 * ps + srow;
            dst[ps] = (valid && rr < p.rows)
                          ? *reinterpret_cast<const float4*>(addb + static_cast<size_t>(rr) * p.n + c * 32 + 4 * sslot)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        };
        tmem_ld32(t_acc, v[0]);
        load_add(0, add[0]);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          tmem_wait_ld();
          if (c + 1 < NC) {
            tmem_ld32(t_acc + (c + 1) * 32, v[(c + 1) & 1]);
            load_add(c + 1, add[(c + 1) & 1]);
          } else {
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * as);
          }
          const uint32_t(&vc)[32] = v[c & 1];
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = (p.bias != nullptr) ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            float a0 = __uint_as_float(vc[4 * j]) + bv.x, a1 = __uint_as_float(vc[4 * j + 1]) + bv.y;
            float a2 = __uint_as_float(vc[4 * j + 2]) + bv.z, a3 = __uint_as_float(vc[4 * j + 3]) + bv.w;
            if constexpr (EPI == EPI_BIAS_GELU_POS_F32) {
              a0 = gelu_erf_fast(a0); a1 = gelu_erf_fast(a1);
              a2 = gelu_erf_fast(a2); a3 = gelu_erf_fast(a3);
            }
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stg_row + ((j ^ (lane & 7)) << 4)),
                         "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int ps = 0; ps < 8; ++ps) {
            const int row = 4 * ps + srow;
            float4 d = *reinterpret_cast<const float4*>(stg + row * 128 + ((sslot ^ (row & 7)) << 4));
            const float4 ad = add[c & 1][ps];
            d.x += ad.x; d.y += ad.y; d.z += ad.z; d.w += ad.w;
            const int rr = rbase + row;
            if (valid && rr < p.rows)
              *reinterpret_cast<float4*>(cb + static_cast<size_t>(rr) * p.c_row_stride + c * 32 + 4 * sslot) = d;
          }
          __syncwarp();
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace gww
