// Q-Adapter convolutions on the 5th-generation tensor cores.
//
// conv2 (3x3, 16->32, 256x256) and conv3 (3x3, 32->64, 128x128) of the reference's freq_adapter
// (MLGWSC-1/inference.py:322-333) are 604 MFLOP each per det-window -- 94 % of the Q-Adapter -- and ran as fp32
// CUDA-core kernels at ~46 TFLOP/s in round 1 (56.8 ms of the 150 ms MLGWSC-1 step).  Here they are implicit
// GEMMs on tcgen05.mma:
//
//   * M = 128 output pixels (a tile of 16 rows x 8 columns), N = Cout (32 / 64), K = 9 taps x Cin.
//   * precision: the 1e-4 feature gate rules out plain bf16 / tf32.  Every fp32 value v is split into two bf16
//     numbers hi = bf16(v), lo = bf16(v - hi) (16 significant bits) and the product is formed as
//     hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM, relative error ~1e-5 per product (the dropped lo*lo
//     term is 2^-18), measured 1.3e-5 on the features (tests).  The three terms take TWO MMAs per (tap, 16-channel
//     K step): A_hi x [W_hi | W_lo] with the hi and lo weights concatenated along N (two accumulator column blocks,
//     summed in the epilogue) and A_lo x W_hi into the first block.  These small-N MMAs are bound by the tensor
//     core's shared-memory operand fetch (a 4 KB A tile per MMA against 16-32 clocks of math), so one A read less
//     per step is worth 25 % (r2 launch list: conv2 was at 2.7x its math time).
//   * no im2col: the haloed input tile (18 x 10 pixels) is staged once in shared memory as PLANES of 16-byte
//     elements -- plane p holds 8 consecutive channels (hi or lo) of every pixel, pixel-major.  In the
//     no-swizzle K-major canonical layout of the UMMA shared-memory descriptor (8 rows x 16 B core matrices,
//     rows 16 B apart, 8-row groups SBO apart, the two K halves LBO apart) a tile row of 8 pixels IS a core
//     matrix, the next tile row is SBO = 10 x 16 B further, the next 8 channels are LBO = one plane further, and
//     a tap (dy, dx) is nothing but the descriptor's start address moved by (dy*10 + dx) x 16 B.  Nine taps =
//     nine descriptors into the same staged tile; zero padding comes from the staging loads.
//   * weights live in shared memory in the same canonical layout ([tap][plane][Cout] x 16 B, hi and lo).
//   * one persistent CTA per SM made of G independent 128-thread groups (named barriers): each group runs a
//     software pipeline over its tiles with two staging buffers -- the TMA box of tile t+2 in flight, tile t+1
//     resident, 18 / 36 MMAs of tile t issued by one elected thread -> tcgen05.commit -> tcgen05.ld -> epilogue of
//     tile t overlapping the MMAs of tile t+1 -- and the groups interleave on the tensor pipe.  Activations move
//     between the kernels in the plane format (bf16 hi/lo): act1 4 planes of 256x256, act2 8 planes of 128x128 --
//     the same bytes as the fp32 NHWC tensors of round 1.
//   * epilogues: conv2 = 2x2 max-pool by warp shuffles (the 16x8 tile maps pool partners to lanes ^1 and ^8),
//     bias, ReLU, hi/lo split, 16-byte plane stores; conv3 = bias, ReLU, the 1x1 conv (64 -> 1) as a dot product.
#pragma once
#include "ptx.cuh"

namespace gww {

constexpr int kQtTileW = 8, kQtTileH = 16;               // output pixels per tile (M = 128)
constexpr int kQtHaloW = kQtTileW + 2, kQtHaloH = kQtTileH + 2;
constexpr int kQtPlaneElems = kQtHaloW * kQtHaloH;      // 180 x 16 B per plane of a staged tile
constexpr int kQtPlaneBytes = kQtPlaneElems * 16;       // 2880
constexpr int kQtGroupThreads = 128;

template <int CIN, int COUT>
struct QtCfg {
  static constexpr int kChunks = CIN / 8;                // 16-byte channel chunks per pixel (2 / 4)
  static constexpr int kPlanes = 2 * kChunks;            // hi planes then lo planes
  static constexpr int kTileBytes = kPlanes * kQtPlaneBytes;
  static constexpr int kWElems = 9 * kPlanes * COUT;     // 16-byte weight elements
  static constexpr int kWBytes = kWElems * 16;
  static constexpr int kGroups = (CIN == 16) ? 6 : 3;     // shared memory (two staging buffers per group) and registers
  static constexpr int kBufs = 2;                         // staging buffers per group: tile t+2 loads while t+1 computes
  static constexpr int kThreads = kGroups * kQtGroupThreads;
  static constexpr int kAccCols = 2 * COUT;               // [A_hi W_hi + A_lo W_hi | A_hi W_lo]
  static constexpr int kTmemCols = (kGroups * kAccCols <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kWBytes + kGroups * kBufs * kTileBytes + COUT * 8 + 64 + kGroups * 8 * (1 + kBufs) + 16;
};

// no-swizzle K-major shared-memory matrix descriptor (see header comment)
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;                   // descriptor version (sm_100)
  return d;                                              // layout type 0 = SWIZZLE_NONE
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}


__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 ah = __float2bfloat16(a), bh = __float2bfloat16(b);
  const __nv_bfloat16 al = __float2bfloat16(a - __bfloat162float(ah));
  const __nv_bfloat16 bl = __float2bfloat16(b - __bfloat162float(bh));
  hi = static_cast<uint32_t>(__bfloat16_as_ushort(ah)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bh)) << 16);
  lo = static_cast<uint32_t>(__bfloat16_as_ushort(al)) | (static_cast<uint32_t>(__bfloat16_as_ushort(bl)) << 16);
}

// Weights of a 3x3 convolution, host layout [tap][ci][co] fp32 (as uploaded for the fp32 kernels), packed once
// per model into the shared-memory image: element (((tap * kChunks + chunk) * 2 + hl) * COUT + co) holds channels
// 8*chunk .. 8*chunk+7 of output channel co as bf16 (hl = 0: hi, 1: lo), i.e. per (tap, chunk) a [2*COUT] x 16 B
// K-major operand whose first COUT rows are W_hi and whose last COUT rows are W_lo.
template <int CIN, int COUT>
__global__ void qt_pack_weights_kernel(const float* __restrict__ w, uint4* __restrict__ out) {
  using Cfg = QtCfg<CIN, COUT>;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cfg::kWElems) return;
  const int co = i % COUT, hl = (i / COUT) & 1, chunk = (i / (2 * COUT)) % Cfg::kChunks, tap = i / (2 * COUT * Cfg::kChunks);
  const bool want_lo = hl != 0;
  uint32_t r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float a = w[(tap * CIN + 8 * chunk + 2 * e) * COUT + co];
    const float b = w[(tap * CIN + 8 * chunk + 2 * e + 1) * COUT + co];
    uint32_t hi, lo;
    split_bf16x2(a, b, hi, lo);
    r[e] = want_lo ? lo : hi;
  }
  out[i] = make_uint4(r[0], r[1], r[2], r[3]);
}

// ------------------------------------------------------------------------------------------------
// conv3x3(CIN -> COUT, pad 1) on plane-format input [n][kPlanes][H][W] x 16 B.
//   MODE 0 (conv2): + bias + ReLU + maxpool2 -> plane-format output [n][2*COUT/8][H/2][W/2] x 16 B
//   MODE 1 (conv3): + bias + ReLU + conv1x1(COUT -> 1) + b4 -> map [n][H][W] f32
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(QtCfg<CIN, COUT>::kThreads, 1)
qadapter_conv_tc_kernel(const __grid_constant__ CUtensorMap tmIn,   // {4 W (u32), H, kPlanes, n} box {40, 18, kPlanes, 1}
                        const uint4* __restrict__ wpacked, const float* __restrict__ bias,
                        const float* __restrict__ w4, float b4, void* __restrict__ out, int H, int W, long n_img) {
  using Cfg = QtCfg<CIN, COUT>;
  extern __shared__ __align__(128) uint8_t qt_smem[];
  uint4* w_s = reinterpret_cast<uint4*>(qt_smem);
  uint8_t* tiles = qt_smem + Cfg::kWBytes;
  float* bias_s = reinterpret_cast<float*>(tiles + Cfg::kGroups * Cfg::kBufs * Cfg::kTileBytes);
  float* w4_s = bias_s + COUT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w4_s + COUT + 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + Cfg::kGroups * (1 + Cfg::kBufs));   // [grp]: MMAs done; then [grp][buf]: tile landed

  const int tid = threadIdx.x;
  const int grp = uniform_warp_idx() >> 2;               // 4 warps per group
  const int gt = tid & (kQtGroupThreads - 1);            // thread within the group
  const int wig = (tid >> 5) & 3;                        // warp within the group == TMEM lane quarter

  for (int i = tid; i < Cfg::kWElems; i += Cfg::kThreads) w_s[i] = wpacked[i];
  for (int i = tid; i < COUT; i += Cfg::kThreads) {
    bias_s[i] = bias[i];
    w4_s[i] = (MODE == 1) ? w4[i] : 0.f;
  }
  if (tid < Cfg::kGroups * (1 + Cfg::kBufs)) mbar_init(smem_u32(&bars[tid]), 1);
  if (tid == 0) tma_prefetch_desc(&tmIn);
  if (uniform_warp_idx() == 0) {
    tmem_alloc<Cfg::kTmemCols>(smem_u32(tmem_slot));
    tmem_relinquish();
  }
  fence_mbar_init();
  fence_proxy_async_smem();                              // weights were written with generic stores
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(grp * Cfg::kAccCols);   // this group's accumulator columns
  const uint32_t bar = smem_u32(&bars[grp]);
  const uint32_t bar_full0 = smem_u32(&bars[Cfg::kGroups + grp * Cfg::kBufs]);                // + 8 * buffer
  const uint32_t tile_addr0 = smem_u32(tiles + grp * Cfg::kBufs * Cfg::kTileBytes);   // buffer b at + b * kTileBytes
  const uint32_t w_addr = smem_u32(w_s);

  const int tiles_x = W / kQtTileW, tiles_y = H / kQtTileH;
  const long tiles_per_img = static_cast<long>(tiles_x) * tiles_y;
  const long n_tiles = tiles_per_img * n_img;
  const long stride = static_cast<long>(gridDim.x) * Cfg::kGroups;
  constexpr uint32_t kIdescWide = make_idesc_bf16(128, 2 * COUT), kIdescNarrow = make_idesc_bf16(128, COUT);
  auto issue_mmas = [&](const uint32_t tile_addr) {
    // per tap and pair of 8-channel chunks (K = 16):  D[:, 0:2C] (+)= A_hi x [W_hi | W_lo] ;  D[:, 0:C] += A_lo x W_hi
    uint32_t first = 1;
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const uint32_t shift = static_cast<uint32_t>(((tap / 3) * kQtHaloW + (tap % 3)) * 16);
#pragma unroll
      for (int kp = 0; kp < Cfg::kChunks / 2; ++kp) {
        const uint64_t a_hi = make_nosw_desc(tile_addr + (2 * kp) * kQtPlaneBytes + shift, kQtPlaneBytes, kQtHaloW * 16);
        const uint64_t a_lo = make_nosw_desc(tile_addr + (Cfg::kChunks + 2 * kp) * kQtPlaneBytes + shift, kQtPlaneBytes, kQtHaloW * 16);
        const uint64_t bdesc = make_nosw_desc(w_addr + ((tap * Cfg::kChunks + 2 * kp) * 2 * COUT) * 16, 2 * COUT * 16, 128);
        umma_ss(tmem_acc, a_hi, bdesc, kIdescWide, first ? 0u : 1u);
        umma_ss(tmem_acc, a_lo, bdesc, kIdescNarrow, 1u);
        first = 0;
      }
    }
  };
  // stage the haloed tile of tile index t into buffer b: ONE TMA box (4-D: 16-byte pixels as 4 words, rows, planes,
  // image; out-of-image coordinates are zero-filled = the convolution's padding).  r2 ncu of the cp.async version: 720 /
  // 1440 16-byte copies per tile with their index arithmetic were ~40 % of the kernel's instructions and kept the L1
  // 74 % busy.  Called by one elected thread.
  auto stage_tile = [&](long t, const int b) {
    const long img = t / tiles_per_img;
    const int rem = static_cast<int>(t - img * tiles_per_img);
    const int ty0 = (rem / tiles_x) * kQtTileH - 1, tx0 = (rem % tiles_x) * kQtTileW - 1;
    mbar_arrive_expect_tx(bar_full0 + 8 * b, Cfg::kTileBytes);
    tma_load_4d(tile_addr0 + b * Cfg::kTileBytes, &tmIn, bar_full0 + 8 * b, 4 * tx0, ty0, 0, static_cast<int>(img));
  };
  // accumulator drained by the group's four warps (named barrier) and tile landed (mbarrier) -> one thread issues the MMAs
  auto publish_and_issue = [&](const int b, const uint32_t parity) {
    named_bar_sync(1 + grp, kQtGroupThreads);
    if (wig == 0) {
      mbar_wait(bar_full0 + 8 * b, parity);
      tc_fence_after();
      if (elect_one()) {
        issue_mmas(tile_addr0 + b * Cfg::kTileBytes);
        umma_commit(bar);
      }
      __syncwarp();
    }
  };

  // Software pipeline per group (two staging buffers): while the MMAs of tile t run, tile t+1 is already resident and
  // tile t+2 is in flight (TMA); the epilogue of tile t overlaps the MMAs of tile t+1.  Tile number `it` of the group
  // lives in buffer it & 1, the (it >> 1)-th use of that buffer.
  long t = static_cast<long>(blockIdx.x) * Cfg::kGroups + grp;
  uint32_t phase = 0;
  int it = 0;
  if (t < n_tiles) {
    if (wig == 0) {
      if (elect_one()) {
        stage_tile(t, 0);
        if (t + stride < n_tiles) stage_tile(t + stride, 1);
      }
      __syncwarp();
    }
    publish_and_issue(0, 0);
  }
  while (t < n_tiles) {
    const long tn = t + stride, tnn = tn + stride;
    mbar_wait(bar, phase);                               // MMAs of tile t done: accumulator ready, buffer of t free
    phase ^= 1;
    tc_fence_after();
    if (wig == 0 && tnn < n_tiles) {                     // two tiles ahead, into the buffer tile t just released
      if (elect_one()) stage_tile(tnn, it & 1);
      __syncwarp();
    }
    uint32_t acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT / 32; ++c) {
      uint32_t lo_blk[32];
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(wig * 32) << 16) + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32 * c]));
      tmem_ld32(tmem_acc + (static_cast<uint32_t>(wig * 32) << 16) + COUT + c * 32, lo_blk);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[32 * c + i] = __float_as_uint(__uint_as_float(acc[32 * c + i]) + __uint_as_float(lo_blk[i]));
    }
    tc_fence_before();
    if (tn < n_tiles) publish_and_issue((it + 1) & 1, static_cast<uint32_t>(((it + 1) >> 1) & 1));
    // ---- epilogue of tile t
    const long img = t / tiles_per_img;
    const int rem = static_cast<int>(t - img * tiles_per_img);
    const int y0 = (rem / tiles_x) * kQtTileH, x0 = (rem % tiles_x) * kQtTileW;
    const int lane = tid & 31;
    const int m = wig * 32 + lane;                       // GEMM row == pixel (m / 8, m % 8) of the tile
    if constexpr (MODE == 0) {
      // 2x2 max-pool: partners are lanes ^1 (x) and ^8 (y).  Each exchange halves the channels a lane keeps, so
      // the four lanes of a quad end with 8 different channels each: chunk q = 2 * (lane & 1) + ((lane >> 3) & 1).
      static_assert(COUT == 32, "conv2 epilogue is written for 32 output channels");
      const bool bx = lane & 1, by = (lane >> 3) & 1;
      float h16[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float mine_keep = __uint_as_float(bx ? acc[16 + i] : acc[i]);
        const float mine_send = __uint_as_float(bx ? acc[i] : acc[16 + i]);
        h16[i] = fmaxf(mine_keep, __shfl_xor_sync(0xffffffffu, mine_send, 1));
      }
      float h8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float keep = by ? h16[8 + i] : h16[i];
        const float send = by ? h16[i] : h16[8 + i];
        h8[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 8));
      }
      const int q = 2 * static_cast<int>(bx) + static_cast<int>(by);   // channels 8q .. 8q+7
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = fmaxf(h8[2 * e] + bias_s[8 * q + 2 * e], 0.f);
        const float b = fmaxf(h8[2 * e + 1] + bias_s[8 * q + 2 * e + 1], 0.f);
        split_bf16x2(a, b, hi[e], lo[e]);
      }
      const int PH = H >> 1, PW = W >> 1;
      const int py = (y0 + (m >> 3)) >> 1, px = (x0 + (m & 7)) >> 1;
      uint4* dst = reinterpret_cast<uint4*>(out) + img * static_cast<long>(COUT / 4) * PH * PW;   // 2*COUT/8 planes
      dst[(static_cast<long>(q) * PH + py) * PW + px] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dst[(static_cast<long>(COUT / 8 + q) * PH + py) * PW + px] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      float s = b4;
#pragma unroll
      for (int c = 0; c < COUT; ++c) s = fmaf(fmaxf(__uint_as_float(acc[c]) + bias_s[c], 0.f), w4_s[c], s);
      float* map = reinterpret_cast<float*>(out) + img * static_cast<long>(H) * W;
      map[static_cast<long>(y0 + (m >> 3)) * W + x0 + (m & 7)] = s;
    }
    t = tn;
    ++it;
  }
  tc_fence_before();
  __syncthreads();
  if (uniform_warp_idx() == 0) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

}  // namespace gww
