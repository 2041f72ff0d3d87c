// Front end B: Q-transform (QScan) + Q-Adapter CNN for the MLGWSC-1 search (sm_100a).
//
// Replaces, per detector and per 256-window batch, the reference's
//   ml4gw.transforms.QScan(duration=1, fs=2048, qrange=[4,128], spectrogram_shape=[512,512])
//                                          MLGWSC-1/inference.py:316-321, called at :345
//   QTransformAdapter.freq_adapter / final_pool / scale,bias / FiLM     inference.py:322-351
// The QScan arithmetic lives in the third-party ml4gw package (absent, unpinned): the specification
// followed here is oracle/qscan.py (GWpy 3.0.8 Q-transform as batched by ml4gw), SURVEY.md section 8c.
//
// Kernels
//   qscan_tiles_kernel   one CTA per det-window: FFT-2048 (norm="forward", one-sided x2) in smem, then
//                        all 148 (q, f) rows of all 5 planes: bisquare-windowed gather, zero-pad,
//                        ifftshift, inverse FFT (128..2048 points), |.|^2, / median.  Rows of <= 512
//                        tiles are processed one per warp (warp-synchronous Stockham FFT, median by
//                        bisection on the float bit pattern, no sort); the 12 longer rows by the CTA.
//                        Writes normalised tile energies [n, 49664] and atomically maxes the
//                        per-plane peak energy (the batch-coupled plane choice of QScan.forward).
//   qscan_interp_kernel  arg-max plane -> cubic-convolution (A=-0.75, align_corners=False) along time
//                        to 512 per row (smem), then along frequency to 512 -> spec [n,512,512] f32.
//   qadapter_conv1_kernel  conv3x3(1->16)+ReLU+maxpool2   spec -> act1 [n,256,256,16]  (NHWC f32)
//   qadapter_conv2_kernel  conv3x3(16->32)+ReLU+maxpool2  act1 -> act2 [n,128,128,32]  (NHWC f32)
//   qadapter_conv3_kernel  conv3x3(32->64)+ReLU+conv1x1(64->1)   act2 -> map [n,128,128] f32
//   qadapter_pool_kernel   AdaptiveAvgPool2d((80,3000)) + scale/bias + FiLM -> f32 [n,80,3000] and/or
//                          bf16 time-major [.,3002,80] (the layout the encoder's conv-stem TMA reads)
// The convolutions are fp32 FFMA register-tiled direct convolutions (the 1e-4 feature gate rules out
// single-pass bf16 tensor-core products; a split-bf16 tcgen05 implicit GEMM is the planned upgrade).
#pragma once
#include "ptx.cuh"

namespace gww {

// ------------------------------------------------------------------------------------------------
// QScan
// ------------------------------------------------------------------------------------------------
struct QRow {
  int n;          // tiles in the row (power of two, 128..2048)
  int log2n;
  int ws;         // bisquare window length (odd)
  int left;       // zeros before the window in the padded array
  int idx0;       // first FFT bin of the window
  int woff;       // offset of the window values in the window table
  int eoff;       // offset of the row's energies inside one det-window's tile buffer
  int plane;
};

struct QPlan {
  const QRow* rows;        // [n_rows] sorted: warp-level rows (n <= 512) first, then CTA-level rows
  const float* window;     // packed window tables
  const float2* tw2048;    // [2048] e^{+2 pi i k / 2048}
  int n_rows;
  int n_rows_warp;         // rows [0, n_rows_warp) have n <= 512
  int n_tiles;             // energies per det-window
  int n_planes;
  int plane_row0[8];       // first row (in ORIGINAL plane order) of each plane in `orig`
  int plane_nrows[8];
  const QRow* orig;        // rows in plane / frequency order (for interpolation)
};

#ifndef GWW_QS_DEBUG_SKIP
#define GWW_QS_DEBUG_SKIP 0     // tuning builds only: 1 = no median (warp rows), 2 = no median (CTA rows), 4 = no CTA rows, 8 = no warp rows
#endif
constexpr int kQsThreads = 512;
constexpr int kQsWarps = kQsThreads / 32;
constexpr int kQsWarpScratch = 2 * 512;                       // float2 elements per warp (ping-pong)
constexpr int kQsSmemBytes = 1032 * 8 + kQsWarps * kQsWarpScratch * 8 + 64 * 4;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// radix-2 Stockham autosort FFT over `n` points in shared memory, executed by `nth` threads whose
// index is `t`; SIGN=+1 inverse (e^{+i}), -1 forward.  Returns the buffer holding the result.
template <int SIGN, typename SyncFn>
__device__ __forceinline__ float2* fft_stockham(float2* a, float2* b, int n, int log2n, int t, int nth,
                                                const float2* __restrict__ tw2048, SyncFn sync) {
  const int half = n >> 1;
  const int tw_stride = 2048 / n;
#pragma unroll 1
  for (int st = 0; st < log2n; ++st) {
    const int s = 1 << st;
    const int m = half >> st;             // half-size of the current sub-transform
    for (int i = t; i < half; i += nth) {
      const int p = i >> st, q = i & (s - 1);
      float2 w = tw2048[(p * s) * tw_stride];
      if (SIGN < 0) w.y = -w.y;
      const float2 u = a[q + s * p], v = a[q + s * (p + m)];
      b[q + s * (2 * p)] = make_float2(u.x + v.x, u.y + v.y);
      b[q + s * (2 * p + 1)] = cmulf(make_float2(u.x - v.x, u.y - v.y), w);
    }
    sync();
    float2* tmp = a; a = b; b = tmp;
  }
  return a;
}

// Radix select, one 8-bit digit per pass (4 passes for non-negative floats compared as uint32): the bin of a 256-entry
// histogram in which the cumulative count first reaches k (1-based), and the count in the bins below it.  Executed by
// every lane of a warp (lane l holds bins 8l .. 8l+7); the result is warp-uniform.  (Rounds 1-2 bisected on the bit
// pattern: 31 count-and-reduce rounds per row -- with one __syncthreads each for the CTA-level rows -- which the r2
// ablation put at 34 % of the whole QScan front end, tools/qscan_bench.py.)
__device__ __forceinline__ void radix_pick_bin(const int* __restrict__ hist, int k, int lane, int& bin, int& below) {
  int c[8], s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; s += c[j]; }
  int incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  const int excl = incl - s;
  const bool mine = (excl < k) && (k <= incl);                 // exactly one lane (k <= total)
  int b = 0, bl = excl;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (bl + c[j] < k && j < 7) { bl += c[j]; b = j + 1; }
    else break;
  }
  const unsigned m = __ballot_sync(0xffffffffu, mine);
  const int src = __ffs(m) - 1;
  bin = __shfl_sync(0xffffffffu, lane * 8 + b, src);
  below = __shfl_sync(0xffffffffu, bl, src);
}

__global__ void __launch_bounds__(kQsThreads, 1)
qscan_tiles_kernel(const float* __restrict__ strain, long n_detwin, long win_stride,
                   float* __restrict__ tiles, unsigned int* __restrict__ plane_max, const QPlan pl) {
  extern __shared__ __align__(16) uint8_t qs_smem[];
  float2* X = reinterpret_cast<float2*>(qs_smem);                  // [1025] one-sided spectrum
  float2* scratch = X + 1032;                                      // kQsWarps x 1024 float2
  int* ctr = reinterpret_cast<int*>(scratch + kQsWarps * kQsWarpScratch);
  int* cnt_s = ctr + 8;                                            // block-wide counters

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long w = blockIdx.x;
  if (w >= n_detwin) return;
  const float* x = strain + w * win_stride;
  float* out = tiles + w * static_cast<long>(pl.n_tiles);
  auto bsync = [] { __syncthreads(); };
  auto wsync = [] { __syncwarp(); };

  // ---- X = rfft(x) / 2048, X[1:] *= 2
  {
    float2* fa = scratch;
    float2* fb = scratch + 2048;
    for (int i = tid; i < 2048; i += kQsThreads) fa[i] = make_float2(x[i], 0.f);
    __syncthreads();
    float2* r = fft_stockham<-1>(fa, fb, 2048, 11, tid, kQsThreads, pl.tw2048, bsync);
    for (int i = tid; i <= 1024; i += kQsThreads) {
      const float sc = (i == 0 ? 1.0f : 2.0f) / 2048.0f;
      X[i] = make_float2(r[i].x * sc, r[i].y * sc);
    }
    if (tid == 0) ctr[0] = 0;
    __syncthreads();
  }

  float pmax[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) pmax[i] = 0.f;

  // ---- rows with n <= 512: one warp per row, dynamic assignment
  {
    float2* a = scratch + warp * kQsWarpScratch;
    float2* b = a + 512;
    for (;;) {
      int ri = 0;
      if (lane == 0) ri = atomicAdd(&ctr[0], 1);
      ri = __shfl_sync(0xffffffffu, ri, 0);
      if (ri >= pl.n_rows_warp || (GWW_QS_DEBUG_SKIP & 8)) break;
      const QRow row = pl.rows[ri];
      const int n = row.n;
      for (int k = lane; k < n; k += 32) {
        const int j = (k + (n >> 1)) & (n - 1);
        const int i = j - row.left;
        float2 v = make_float2(0.f, 0.f);
        if (i >= 0 && i < row.ws) {
          const float wv = __ldg(pl.window + row.woff + i);
          const float2 xv = X[min(row.idx0 + i, 1024)];
          v = make_float2(xv.x * wv, xv.y * wv);
        }
        a[k] = v;
      }
      __syncwarp();
      float2* r = fft_stockham<1>(a, b, n, row.log2n, lane, 32, pl.tw2048, wsync);
      // energies into registers (<= 16 per lane)
      const float inv_n = 1.0f / static_cast<float>(n);
      float e[16];
      const int per = n >> 5;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < per) {
          const float2 c = r[lane + 32 * i];
          const float re = c.x * inv_n, im = c.y * inv_n;
          e[i] = re * re + im * im;
        } else {
          e[i] = 0.f;
        }
      }
      const int k1 = n >> 1;                                 // sorted[n/2 - 1] (0-based) is the k1-th smallest
      // radix select over the row's n energies; the FFT buffers are dead (energies are in registers): histogram there
      int* hist = reinterpret_cast<int*>(a);
      uint32_t lo_bits = 0u;
      int kk = k1, n_le = 0;
      __syncwarp();
#pragma unroll 1
      for (int pass = 0; pass < 4 && !(GWW_QS_DEBUG_SKIP & 1); ++pass) {
        const int shift = 24 - 8 * pass;
#pragma unroll
        for (int j = 0; j < 8; ++j) hist[lane * 8 + j] = 0;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i < per) {
            const uint32_t v = __float_as_uint(e[i]);
            if (pass == 0 || (v >> (shift + 8)) == (lo_bits >> (shift + 8))) atomicAdd(&hist[(v >> shift) & 255u], 1);
          }
        }
        __syncwarp();
        int bin, below;
        radix_pick_bin(hist, kk, lane, bin, below);
        if (pass == 3) n_le = (k1 - kk) + below + hist[bin];   // elements <= the selected value
        kk -= below;
        lo_bits |= static_cast<uint32_t>(bin) << shift;
        __syncwarp();
      }
      if (GWW_QS_DEBUG_SKIP & 1) { lo_bits = 0x3f800000u; n_le = n; }
      uint32_t hi_bits = lo_bits;
      if (n_le < k1 + 1) {                                   // next distinct value above
        uint32_t mn = 0x7f800000u;
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < per && __float_as_uint(e[i]) > lo_bits) mn = min(mn, __float_as_uint(e[i]));
        hi_bits = __reduce_min_sync(0xffffffffu, mn);
      }
      const float lo_v = __uint_as_float(lo_bits), hi_v = __uint_as_float(hi_bits);
      const float med = lo_v + 0.5f * (hi_v - lo_v);         // torch.quantile(q=0.5), linear
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < per) {
          const float v = e[i] / med;
          out[row.eoff + lane + 32 * i] = v;
          mx = fmaxf(mx, v);
        }
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p == row.plane) pmax[p] = fmaxf(pmax[p], mx);
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- rows with n >= 1024: the whole CTA per row
  {
    float2* a = scratch;
    float2* b = scratch + 2048;
    for (int ri = pl.n_rows_warp; ri < ((GWW_QS_DEBUG_SKIP & 4) ? 0 : pl.n_rows); ++ri) {
      const QRow row = pl.rows[ri];
      const int n = row.n;
      for (int k = tid; k < n; k += kQsThreads) {
        const int j = (k + (n >> 1)) & (n - 1);
        const int i = j - row.left;
        float2 v = make_float2(0.f, 0.f);
        if (i >= 0 && i < row.ws) {
          const float wv = __ldg(pl.window + row.woff + i);
          const float2 xv = X[min(row.idx0 + i, 1024)];
          v = make_float2(xv.x * wv, xv.y * wv);
        }
        a[k] = v;
      }
      __syncthreads();
      float2* r = fft_stockham<1>(a, b, n, row.log2n, tid, kQsThreads, pl.tw2048, bsync);
      const float inv_n = 1.0f / static_cast<float>(n);
      float e[4];
      const int per = n / kQsThreads;                        // 2 or 4
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < per) {
          const float2 c = r[tid + kQsThreads * i];
          const float re = c.x * inv_n, im = c.y * inv_n;
          e[i] = re * re + im * im;
        } else {
          e[i] = 0.f;
        }
      }
      __syncthreads();                                       // r (== a or b) is dead: next row may overwrite it
      // block-wide radix select: four 256-bin histograms (one per pass, cleared together) in the dead FFT buffer; one
      // barrier per pass, every warp scans the histogram for itself
      int* hist = reinterpret_cast<int*>(scratch);           // [4][256]
      for (int i = tid; i < 4 * 256; i += kQsThreads) hist[i] = 0;
      __syncthreads();
      const int k1 = n >> 1;
      uint32_t lo_bits = 0u;
      int kk = k1, n_le = 0;
#pragma unroll 1
      for (int pass = 0; pass < 4 && !(GWW_QS_DEBUG_SKIP & 2); ++pass) {
        const int shift = 24 - 8 * pass;
        int* h = hist + 256 * pass;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i < per) {
            const uint32_t v = __float_as_uint(e[i]);
            if (pass == 0 || (v >> (shift + 8)) == (lo_bits >> (shift + 8))) atomicAdd(&h[(v >> shift) & 255u], 1);
          }
        }
        __syncthreads();
        int bin, below;
        radix_pick_bin(h, kk, lane, bin, below);
        if (pass == 3) n_le = (k1 - kk) + below + h[bin];
        kk -= below;
        lo_bits |= static_cast<uint32_t>(bin) << shift;
      }
      if (GWW_QS_DEBUG_SKIP & 2) { lo_bits = 0x3f800000u; n_le = n; }
      uint32_t hi_bits = lo_bits;
      if (n_le < k1 + 1) {
        uint32_t mn = 0x7f800000u;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < per && __float_as_uint(e[i]) > lo_bits) mn = min(mn, __float_as_uint(e[i]));
        mn = __reduce_min_sync(0xffffffffu, mn);
        unsigned int* mslot = reinterpret_cast<unsigned int*>(cnt_s + 4);
        if (tid == 0) *mslot = 0x7f800000u;
        __syncthreads();
        if (lane == 0) atomicMin(mslot, mn);
        __syncthreads();
        hi_bits = *mslot;
      }
      const float lo_v = __uint_as_float(lo_bits), hi_v = __uint_as_float(hi_bits);
      const float med = lo_v + 0.5f * (hi_v - lo_v);
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < per) {
          const float v = e[i] / med;
          out[row.eoff + tid + kQsThreads * i] = v;
          mx = fmaxf(mx, v);
        }
      }
#pragma unroll
      for (int p = 0; p < 8; ++p) if (p == row.plane) pmax[p] = fmaxf(pmax[p], mx);
      __syncthreads();
    }
  }
  // ---- per-plane peak energy of this det-window -> global (energies >= 0: uint order == float order)
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    float v = pmax[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0 && p < pl.n_planes && v > 0.f) atomicMax(plane_max + p, __float_as_uint(v));
  }
}

// PyTorch upsample_bicubic2d coefficients (A = -0.75), UpSample.h get_cubic_upsample_coefficients
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  c[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  c[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  c[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  c[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}

constexpr int kQiThreads = 256;
constexpr int kQiMaxRows = 48;
constexpr int kQiSplit = 4;      // CTAs per det-window along the output frequency axis

// spec[w, fo, to] from the tile energies of the arg-max plane.  out_f == out_t == 512 supported
// generally as (out_f, out_t) <= (512, 512) with out_t == 512 smem rows.
__global__ void __launch_bounds__(kQiThreads)
qscan_interp_kernel(const float* __restrict__ tiles, const unsigned int* __restrict__ plane_max,
                    float* __restrict__ spec, int* __restrict__ plane_out, int out_f, int out_t,
                    const QPlan pl) {
  extern __shared__ __align__(16) float qi_rows[];     // [R][out_t]
  const long w = blockIdx.x;
  // torch.argmax over planes (first maximum wins)
  int plane = 0;
  unsigned int best = plane_max[0];
  for (int p = 1; p < pl.n_planes; ++p) {
    const unsigned int v = plane_max[p];
    if (v > best) { best = v; plane = p; }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && plane_out != nullptr) *plane_out = plane;
  const int R = pl.plane_nrows[plane];
  const QRow* rows = pl.orig + pl.plane_row0[plane];
  const float* src = tiles + w * static_cast<long>(pl.n_tiles);
  // ---- this CTA's slice of the output frequency axis and the rows of the plane its 4-tap stencils touch
  const int f_per = (out_f + kQiSplit - 1) / kQiSplit;
  const int f0 = blockIdx.y * f_per;
  const int f1 = min(out_f, f0 + f_per);
  const float scale_f = static_cast<float>(R) / static_cast<float>(out_f);
  const int r_lo = max(0, static_cast<int>(floorf(scale_f * (static_cast<float>(f0) + 0.5f) - 0.5f)) - 1);
  const int r_hi = min(R - 1, static_cast<int>(floorf(scale_f * (static_cast<float>(f1 - 1) + 0.5f) - 0.5f)) + 2);
  const int n_r = (f1 > f0) ? r_hi - r_lo + 1 : 0;
  // ---- time axis: those rows -> out_t samples (rounds 1-2: every CTA interpolated all R rows, 4x the work)
  for (int idx = threadIdx.x; idx < n_r * out_t; idx += kQiThreads) {
    const int rl = idx / out_t, t = idx - rl * out_t;
    const QRow row = rows[r_lo + rl];
    const float scale = static_cast<float>(row.n) / static_cast<float>(out_t);
    const float s = scale * (static_cast<float>(t) + 0.5f) - 0.5f;
    const float fl = floorf(s);
    const int ix = static_cast<int>(fl);
    float c[4];
    cubic_coeffs(s - fl, c);
    const float* e = src + row.eoff;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = min(max(ix - 1 + k, 0), row.n - 1);
      acc = fmaf(c[k], __ldg(e + j), acc);
    }
    qi_rows[idx] = acc;
  }
  __syncthreads();
  // ---- frequency axis
  float* dst = spec + w * static_cast<long>(out_f) * out_t;
  for (int fo = f0; fo < f1; ++fo) {
    const float s = scale_f * (static_cast<float>(fo) + 0.5f) - 0.5f;
    const float fl = floorf(s);
    const int iy = static_cast<int>(fl);
    float c[4];
    cubic_coeffs(s - fl, c);
    int rr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) rr[k] = min(max(iy - 1 + k, 0), R - 1) - r_lo;   // clamped rows stay inside [r_lo, r_hi]
    for (int t = threadIdx.x; t < out_t; t += kQiThreads) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc = fmaf(c[k], qi_rows[rr[k] * out_t + t], acc);
      dst[static_cast<long>(fo) * out_t + t] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Q-Adapter CNN (MLGWSC-1/inference.py:322-351)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo);   // qadapter_tc.cuh

struct QAdapterDev {
  const float* w1;   // [9][16]          conv1 weight, tap-major
  const float* b1;   // [16]
  const float* w2;   // [9][16][32]      conv2 weight [tap][ci][co]
  const float* b2;   // [32]
  const float* w3;   // [9][32][64]
  const float* b3;   // [64]
  const float* w4;   // [64]             conv 1x1
  const uint4* w2p;  // conv2 / conv3 weights packed for the tensor-core path (qadapter_tc.cuh), or nullptr
  const uint4* w3p;
  float b4;
  float scale, bias;
  float gamma[8], beta[8];
};

// conv3x3(1->16, pad 1) + ReLU + maxpool2 : spec [n,H,W] -> act1 [n,H/2,W/2,16] (NHWC)
// CTA = 16x16 pooled pixels (32x32 conv pixels); one thread per pooled pixel, 16 channels.
// PLANES: store act1 in the plane format of the tensor-core convolutions (qadapter_tc.cuh): [n][4][H/2][W/2] x 16 B,
// planes = bf16 hi of channels 0-7, hi 8-15, lo 0-7, lo 8-15 (lo = bf16(v - hi)).
template <bool PLANES>
__global__ void __launch_bounds__(256, 4)
qadapter_conv1_kernel(const float* __restrict__ spec, float* __restrict__ act1, int H, int W,
                      const QAdapterDev ad) {
  __shared__ float tile[34][35];
  __shared__ __align__(16) float ws[9 * 16 + 16];
  const long n = blockIdx.z;
  const int py0 = blockIdx.y * 16, px0 = blockIdx.x * 16;
  const float* src = spec + n * static_cast<long>(H) * W;
  for (int i = threadIdx.x; i < 9 * 16; i += 256) ws[i] = ad.w1[i];
  if (threadIdx.x < 16) ws[144 + threadIdx.x] = ad.b1[threadIdx.x];
  for (int i = threadIdx.x; i < 34 * 34; i += 256) {
    const int ty = i / 34, tx = i - ty * 34;
    const int y = 2 * py0 - 1 + ty, x = 2 * px0 - 1 + tx;
    tile[ty][tx] = (y >= 0 && y < H && x >= 0 && x < W) ? src[static_cast<long>(y) * W + x] : 0.f;
  }
  __syncthreads();
  const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
  float p[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) p[a][b] = tile[2 * ly + a][2 * lx + b];
  // 4 conv pixels x 16 channels as packed fp32 channel PAIRS (FFMA2); the nine taps' weights are fetched as
  // four 16-byte broadcasts per tap instead of one 4-byte shared load per multiply-add (r1: the scalar form was
  // bound by those 576 loads per thread: 8.6 ms per MLGWSC step against ~3 ms of arithmetic).  Every
  // accumulator still sees bias, then the taps in (ky, kx) order: bit-identical results.
  // (two passes of 8 channels keep the kernel at <= 64 registers: four CTAs per SM hide the tile-load latency
  //  that bounds this otherwise tiny kernel)
  float o[16];
#pragma unroll
  for (int hc = 0; hc < 2; ++hc) {
    uint64_t acc[4][4];
#pragma unroll
    for (int cp = 0; cp < 4; ++cp) {
      const uint64_t bv = f2_pack(ws[144 + 8 * hc + 2 * cp], ws[144 + 8 * hc + 2 * cp + 1]);
#pragma unroll
      for (int px = 0; px < 4; ++px) acc[px][cp] = bv;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        uint64_t w2[4];
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[(ky * 3 + kx) * 16 + 8 * hc + 4 * c4]);
          w2[2 * c4] = f2_pack(wv.x, wv.y);
          w2[2 * c4 + 1] = f2_pack(wv.z, wv.w);
        }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float iv = p[dy + ky][dx + kx];
            const uint64_t iv2 = f2_pack(iv, iv);
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) acc[dy * 2 + dx][cp] = f2_fma(iv2, w2[cp], acc[dy * 2 + dx][cp]);
          }
      }
#pragma unroll
    for (int cp = 0; cp < 4; ++cp) {
      float b0 = 0.f, b1 = 0.f;     // ReLU floor: max(relu(a), relu(b), ...) == max(0, a, b, ...)
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        float a0, a1;
        f2_unpack(acc[px][cp], a0, a1);
        b0 = fmaxf(b0, a0);
        b1 = fmaxf(b1, a1);
      }
      o[8 * hc + 2 * cp] = b0;
      o[8 * hc + 2 * cp + 1] = b1;
    }
  }
  const int PH = H >> 1, PW = W >> 1;
  if constexpr (PLANES) {
    uint4* dst = reinterpret_cast<uint4*>(act1) + n * 4L * PH * PW + static_cast<long>(py0 + ly) * PW + (px0 + lx);
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16x2(o[8 * hc + 2 * e], o[8 * hc + 2 * e + 1], hi[e], lo[e]);
      dst[static_cast<long>(hc) * PH * PW] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dst[static_cast<long>(2 + hc) * PH * PW] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(act1 + ((n * PH + (py0 + ly)) * static_cast<long>(PW) + (px0 + lx)) * 16);
#pragma unroll
    for (int c = 0; c < 4; ++c) dst[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
  }
}

// The same convolution for the tensor-core path, persistent and TMA-fed: r2 ncu of qadapter_conv1_kernel<true> showed
// 37 % of the warp samples waiting on the global loads of the next 34 x 34 input tile (long scoreboard) although four
// CTAs per SM overlap.  Here CTAs loop over tiles; one thread fetches the NEXT tile with a single TMA box
// ({40, 34} floats starting 4 columns left of the tile: 16-byte aligned rows; out-of-image coordinates are zero-filled = the padding) into the other half
// of a double buffer while the 256 threads compute the current one.  Arithmetic and its order are those of
// qadapter_conv1_kernel<true>: bit-identical outputs.
constexpr int kC1Pitch = 40;   // staged row: image columns 32 bx - 4 .. 32 bx + 35 (the TMA start must be 16-byte aligned)
__global__ void __launch_bounds__(256, 4)
qadapter_conv1_tma_kernel(const __grid_constant__ CUtensorMap tmSpec,   // {W, H, n} f32, box {40, 34, 1}
                          float* __restrict__ act1, int H, int W, long n_img, const QAdapterDev ad) {
  __shared__ __align__(128) float tile[2][1376];      // 34 rows x 40 floats, padded to a 128-byte multiple (TMA destination)
  __shared__ __align__(16) float ws[9 * 16 + 16];
  __shared__ __align__(8) uint64_t full[2];
  for (int i = threadIdx.x; i < 9 * 16; i += 256) ws[i] = ad.w1[i];
  if (threadIdx.x < 16) ws[144 + threadIdx.x] = ad.b1[threadIdx.x];
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmSpec);
    mbar_init(smem_u32(&full[0]), 1);
    mbar_init(smem_u32(&full[1]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int tiles_x = W >> 5, tiles_y = H >> 5;
  const long tiles_per_img = static_cast<long>(tiles_x) * tiles_y;
  const long n_tiles = tiles_per_img * n_img;
  auto stage = [&](long t, int b) {                 // one thread
    const long img = t / tiles_per_img;
    const int rem = static_cast<int>(t - img * tiles_per_img);
    const int by = rem / tiles_x, bx = rem - by * tiles_x;
    mbar_arrive_expect_tx(smem_u32(&full[b]), 34 * kC1Pitch * 4);
    tma_load_3d(smem_u32(&tile[b][0]), &tmSpec, smem_u32(&full[b]), 32 * bx - 4, 32 * by - 1, static_cast<int>(img));
  };
  const int ly = threadIdx.x >> 4, lx = threadIdx.x & 15;
  const int PH = H >> 1, PW = W >> 1;
  long t = blockIdx.x;
  if (t < n_tiles && threadIdx.x == 0) stage(t, 0);
  for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
    const int b = it & 1;
    __syncthreads();                                 // every thread is done with the other buffer (tile it - 1)
    if (threadIdx.x == 0 && t + gridDim.x < n_tiles) stage(t + gridDim.x, b ^ 1);
    mbar_wait(smem_u32(&full[b]), (it >> 1) & 1);
    float p[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float* row = &tile[b][(2 * ly + a) * kC1Pitch + 2 * lx + 3];
      p[a][0] = row[0]; p[a][1] = row[1]; p[a][2] = row[2]; p[a][3] = row[3];
    }
    float o[16];
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      uint64_t acc[4][4];
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        const uint64_t bv = f2_pack(ws[144 + 8 * hc + 2 * cp], ws[144 + 8 * hc + 2 * cp + 1]);
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[px][cp] = bv;
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          uint64_t w2[4];
#pragma unroll
          for (int c4 = 0; c4 < 2; ++c4) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[(ky * 3 + kx) * 16 + 8 * hc + 4 * c4]);
            w2[2 * c4] = f2_pack(wv.x, wv.y);
            w2[2 * c4 + 1] = f2_pack(wv.z, wv.w);
          }
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const float iv = p[dy + ky][dx + kx];
              const uint64_t iv2 = f2_pack(iv, iv);
#pragma unroll
              for (int cp = 0; cp < 4; ++cp) acc[dy * 2 + dx][cp] = f2_fma(iv2, w2[cp], acc[dy * 2 + dx][cp]);
            }
        }
#pragma unroll
      for (int cp = 0; cp < 4; ++cp) {
        float b0 = 0.f, b1 = 0.f;     // ReLU floor: max(relu(a), relu(b), ...) == max(0, a, b, ...)
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          float a0, a1;
          f2_unpack(acc[px][cp], a0, a1);
          b0 = fmaxf(b0, a0);
          b1 = fmaxf(b1, a1);
        }
        o[8 * hc + 2 * cp] = b0;
        o[8 * hc + 2 * cp + 1] = b1;
      }
    }
    const long img = t / tiles_per_img;
    const int rem = static_cast<int>(t - img * tiles_per_img);
    const int py0 = (rem / tiles_x) * 16, px0 = (rem % tiles_x) * 16;
    uint4* dst = reinterpret_cast<uint4*>(act1) + img * 4L * PH * PW + static_cast<long>(py0 + ly) * PW + (px0 + lx);
#pragma unroll
    for (int hc = 0; hc < 2; ++hc) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_bf16x2(o[8 * hc + 2 * e], o[8 * hc + 2 * e + 1], hi[e], lo[e]);
      dst[static_cast<long>(hc) * PH * PW] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      dst[static_cast<long>(2 + hc) * PH * PW] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// conv3x3(16->32, pad 1) + ReLU + maxpool2 : act1 [n,H,W,16] -> act2 [n,H/2,W/2,32]
// CTA = 16x16 conv pixels (8x8 pooled) x 32 output channels.  Thread = (2x2 pixel quad, 8 channels):
// 32 accumulators; input tile in smem channel-major [16][18][24] (pitch 24 makes the float2 loads of a
// half-warp conflict-free), weights [9][16][32] in smem, read as warp-uniform float4 broadcasts.
constexpr int kC2Pitch = 24;
constexpr int kC2SmemBytes = (16 * 18 * kC2Pitch + 9 * 16 * 32 + 32) * 4;
__global__ void __launch_bounds__(256)
qadapter_conv2_kernel(const float* __restrict__ act1, float* __restrict__ act2, int H, int W,
                      const QAdapterDev ad) {
  extern __shared__ __align__(16) float c2_smem[];
  float* tile = c2_smem;                        // [16][18][24]
  float* wsm = tile + 16 * 18 * kC2Pitch;       // [9][16][32]
  float* bsm = wsm + 9 * 16 * 32;               // [32]
  const long n = blockIdx.z;
  const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
  const float* src = act1 + n * static_cast<long>(H) * W * 16;
  for (int i = threadIdx.x; i < 9 * 16 * 32 / 4; i += 256)
    reinterpret_cast<float4*>(wsm)[i] = reinterpret_cast<const float4*>(ad.w2)[i];
  if (threadIdx.x < 32) bsm[threadIdx.x] = ad.b2[threadIdx.x];
  // 18x18 pixels x 16 channels, NHWC in global (64 B per pixel) -> channel-major in smem
  for (int i = threadIdx.x; i < 18 * 18 * 4; i += 256) {
    const int pix = i >> 2, c4 = i & 3;
    const int ty = pix / 18, tx = pix - ty * 18;
    const int y = y0 - 1 + ty, x = x0 - 1 + tx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && x >= 0 && x < W)
      v = *reinterpret_cast<const float4*>(src + (static_cast<long>(y) * W + x) * 16 + 4 * c4);
    float* d = tile + (4 * c4 * 18 + ty) * kC2Pitch + tx;
    d[0] = v.x; d[18 * kC2Pitch] = v.y; d[2 * 18 * kC2Pitch] = v.z; d[3 * 18 * kC2Pitch] = v.w;
  }
  __syncthreads();
  const int cg = threadIdx.x >> 6;              // channel group (8 channels), warp-uniform
  const int quad = threadIdx.x & 63;
  const int qy = quad >> 3, qx = quad & 7;
  // accumulators as packed fp32 PAIRS of adjacent output channels: one FFMA2 (fma.rn.f32x2, per-lane IEEE fma,
  // bit-identical to two fmaf) per two multiply-adds -- the loop is bound by fp32 issue slots (r1: scalar FFMA tops
  // out at 107 of the SM's 128 FMA/clk, FFMA2 reaches 126; tools/ubench/ffma2.cu)
  uint64_t acc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint64_t bv = f2_pack(bsm[cg * 8 + 2 * c], bsm[cg * 8 + 2 * c + 1]);
#pragma unroll
    for (int px = 0; px < 4; ++px) acc[px][c] = bv;
  }
#pragma unroll 2
  for (int ci = 0; ci < 16; ++ci) {
    float p[4][4];
    const float* t = tile + (ci * 18 + 2 * qy) * kC2Pitch + 2 * qx;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float2 v0 = *reinterpret_cast<const float2*>(t + a * kC2Pitch);
      const float2 v1 = *reinterpret_cast<const float2*>(t + a * kC2Pitch + 2);
      p[a][0] = v0.x; p[a][1] = v0.y; p[a][2] = v1.x; p[a][3] = v1.y;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* wp = reinterpret_cast<const float4*>(wsm + ((ky * 3 + kx) * 16 + ci) * 32 + cg * 8);
        const float4 wa = wp[0], wb = wp[1];
        const uint64_t wv[4] = {f2_pack(wa.x, wa.y), f2_pack(wa.z, wa.w), f2_pack(wb.x, wb.y), f2_pack(wb.z, wb.w)};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float iv = p[dy + ky][dx + kx];
            const uint64_t iv2 = f2_pack(iv, iv);
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[dy * 2 + dx][c] = f2_fma(iv2, wv[c], acc[dy * 2 + dx][c]);
          }
      }
  }
  const int PH = H >> 1, PW = W >> 1;
  float o[8];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a0[4], a1[4];
#pragma unroll
    for (int px = 0; px < 4; ++px) f2_unpack(acc[px][c], a0[px], a1[px]);
    o[2 * c] = fmaxf(fmaxf(fmaxf(a0[0], a0[1]), fmaxf(a0[2], a0[3])), 0.f);
    o[2 * c + 1] = fmaxf(fmaxf(fmaxf(a1[0], a1[1]), fmaxf(a1[2], a1[3])), 0.f);
  }
  float4* dst = reinterpret_cast<float4*>(
      act2 + ((n * PH + ((y0 >> 1) + qy)) * static_cast<long>(PW) + ((x0 >> 1) + qx)) * 32 + cg * 8);
  dst[0] = make_float4(o[0], o[1], o[2], o[3]);
  dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// conv3x3(32->64, pad 1) + ReLU + conv1x1(64->1) : act2 [n,H,W,32] -> map [n,H,W]
// CTA = 16x16 pixels x 64 channels; thread = (2x2 quad, 16 channels) -> 64 accumulators; the 32 input
// channels are streamed through smem in two halves of 16.
constexpr int kC3SmemBytes = (16 * 18 * kC2Pitch + 9 * 16 * 64 + 64 + 64 + 4 * 256) * 4;
__global__ void __launch_bounds__(256)
qadapter_conv3_kernel(const float* __restrict__ act2, float* __restrict__ map, int H, int W,
                      const QAdapterDev ad) {
  extern __shared__ __align__(16) float c3_smem[];
  float* tile = c3_smem;                        // [16][18][24]
  float* wsm = tile + 16 * 18 * kC2Pitch;       // [9][16][64]   (current half of ci)
  float* bsm = wsm + 9 * 16 * 64;               // [64]
  float* w4s = bsm + 64;                        // [64]
  float* red = w4s + 64;                        // [4][256]
  const long n = blockIdx.z;
  const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
  const float* src = act2 + n * static_cast<long>(H) * W * 32;
  if (threadIdx.x < 64) { bsm[threadIdx.x] = ad.b3[threadIdx.x]; w4s[threadIdx.x] = ad.w4[threadIdx.x]; }
  const int cg = threadIdx.x >> 6;              // 16 channels per group
  const int quad = threadIdx.x & 63;
  const int qy = quad >> 3, qx = quad & 7;
  uint64_t acc[4][8];                            // packed pairs of adjacent output channels (see conv2)
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint64_t bv = f2_pack(bsm[cg * 16 + 2 * c], bsm[cg * 16 + 2 * c + 1]);
#pragma unroll
    for (int px = 0; px < 4; ++px) acc[px][c] = bv;
  }
  for (int hf = 0; hf < 2; ++hf) {
    __syncthreads();
    // weights of this half: global layout [9][32][64] -> smem [9][16][64]
    for (int i = threadIdx.x; i < 9 * 16 * 64 / 4; i += 256) {
      const int tap = i / (16 * 16), rem = i - tap * (16 * 16);   // rem indexes [16 ci][16 float4]
      reinterpret_cast<float4*>(wsm)[i] =
          reinterpret_cast<const float4*>(ad.w3)[(tap * 32 + hf * 16) * 16 + rem];
    }
    for (int i = threadIdx.x; i < 18 * 18 * 4; i += 256) {
      const int pix = i >> 2, c4 = i & 3;
      const int ty = pix / 18, tx = pix - ty * 18;
      const int y = y0 - 1 + ty, x = x0 - 1 + tx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < H && x >= 0 && x < W)
        v = *reinterpret_cast<const float4*>(src + (static_cast<long>(y) * W + x) * 32 + hf * 16 + 4 * c4);
      float* d = tile + (4 * c4 * 18 + ty) * kC2Pitch + tx;
      d[0] = v.x; d[18 * kC2Pitch] = v.y; d[2 * 18 * kC2Pitch] = v.z; d[3 * 18 * kC2Pitch] = v.w;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < 16; ++ci) {
      float p[4][4];
      const float* t = tile + (ci * 18 + 2 * qy) * kC2Pitch + 2 * qx;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float2 v0 = *reinterpret_cast<const float2*>(t + a * kC2Pitch);
        const float2 v1 = *reinterpret_cast<const float2*>(t + a * kC2Pitch + 2);
        p[a][0] = v0.x; p[a][1] = v0.y; p[a][2] = v1.x; p[a][3] = v1.y;
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(wsm + ((ky * 3 + kx) * 16 + ci) * 64 + cg * 16);
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 wv = wp[c4];
            const uint64_t w01 = f2_pack(wv.x, wv.y), w23 = f2_pack(wv.z, wv.w);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
              for (int dx = 0; dx < 2; ++dx) {
                const float iv = p[dy + ky][dx + kx];
                const uint64_t iv2 = f2_pack(iv, iv);
                uint64_t* a2 = &acc[dy * 2 + dx][2 * c4];
                a2[0] = f2_fma(iv2, w01, a2[0]);
                a2[1] = f2_fma(iv2, w23, a2[1]);
              }
          }
        }
    }
  }
  // ReLU, 1x1 conv partial over this thread's 16 channels, reduce over the 4 channel groups
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float a0, a1;
      f2_unpack(acc[px][c], a0, a1);
      s = fmaf(fmaxf(a0, 0.f), w4s[cg * 16 + 2 * c], s);
      s = fmaf(fmaxf(a1, 0.f), w4s[cg * 16 + 2 * c + 1], s);
    }
    red[cg * 256 + quad * 4 + px] = s;
  }
  __syncthreads();
  {
    const int quad2 = threadIdx.x >> 2, px = threadIdx.x & 3;     // 64 quads x 4 pixels
    const float v = ((red[quad2 * 4 + px] + red[256 + quad2 * 4 + px]) + red[512 + quad2 * 4 + px]) +
                    red[768 + quad2 * 4 + px] + ad.b4;
    const int y = y0 + 2 * (quad2 >> 3) + (px >> 1), x = x0 + 2 * (quad2 & 7) + (px & 1);
    map[(n * H + y) * static_cast<long>(W) + x] = v;
  }
}

// Generic 3x3 convolution (pad 1) + ReLU for adapter geometries other than the inference default (e.g. the
// 128x128 / 32-64-128 adapter of MLGWSC-1/train.py:104,118-123): NHWC fp32 in, weights [9][CI][CO], any CI and
// any CO that is a multiple of 16.  CTA = 16x16 conv pixels, thread = one pixel, 16 output channels at a time.
//   POOL : + maxpool2 -> out NHWC [n, H/2, W/2, CO]
//   FINAL: + conv1x1(CO -> 1) + b4 -> out map [n, H, W]
// Correctness path (fp32 CUDA cores, ~10 TFLOP/s); the default geometry runs on the tensor cores (qadapter_tc.cuh).
template <bool POOL, bool FINAL>
__global__ void __launch_bounds__(256)
qadapter_conv_generic_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                             const float* __restrict__ w4, float b4, float* __restrict__ out, int H, int W, int CI, int CO) {
  extern __shared__ __align__(16) float cg_tile[];          // [18][18][CI]
  const long n = blockIdx.z;
  const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
  const float* src = in + n * static_cast<long>(H) * W * CI;
  for (int i = threadIdx.x; i < 18 * 18 * CI; i += 256) {
    const int pix = i / CI, c = i - pix * CI;
    const int y = y0 - 1 + pix / 18, x = x0 - 1 + pix % 18;
    cg_tile[i] = (y >= 0 && y < H && x >= 0 && x < W) ? src[(static_cast<long>(y) * W + x) * CI + c] : 0.f;
  }
  __syncthreads();
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int y = y0 + ty, x = x0 + tx;
  float fin = b4;
  for (int c0 = 0; c0 < CO; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = __ldg(b + c0 + c);
    for (int tap = 0; tap < 9; ++tap) {
      const float* t = cg_tile + ((ty + tap / 3) * 18 + tx + tap % 3) * CI;
      const float4* wp = reinterpret_cast<const float4*>(w + (static_cast<long>(tap) * CI) * CO + c0);
      for (int ci = 0; ci < CI; ++ci) {
        const float v = t[ci];
        const float4* wr = wp + static_cast<long>(ci) * (CO / 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = __ldg(wr + q);
          acc[4 * q] = fmaf(v, ww.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
        }
      }
    }
    if constexpr (FINAL) {
#pragma unroll
      for (int c = 0; c < 16; ++c) fin = fmaf(fmaxf(acc[c], 0.f), __ldg(w4 + c0 + c), fin);
    } else if constexpr (POOL) {
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float v = fmaxf(acc[c], 0.f);
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));     // x partner
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));    // y partner (a warp holds two tile rows)
        acc[c] = v;
      }
      if (((tx | ty) & 1) == 0 && y < H && x < W) {
        float* dst = out + ((n * (H >> 1) + (y >> 1)) * static_cast<long>(W >> 1) + (x >> 1)) * CO + c0;
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[c] = acc[c];
      }
    } else {
      if (y < H && x < W) {
        float* dst = out + ((n * H + y) * static_cast<long>(W) + x) * CO + c0;
#pragma unroll
        for (int c = 0; c < 16; ++c) dst[c] = fmaxf(acc[c], 0.f);
      }
    }
  }
  if constexpr (FINAL) {
    if (y < H && x < W) out[(n * H + y) * static_cast<long>(W) + x] = fin;
  }
}

// AdaptiveAvgPool2d((OF, OT)) of map [n,H,W] + scale/bias + FiLM(det) -> feats.
//   out_f32: [n, OF, OT] (reference layout) or nullptr
//   out_tm : bf16 time-major, det-window w at out_tm + (w*tm_stride_w + tm_off) * (OT+2)*OF, with zero
//            rows 0 and OT+1 (conv-stem padding), or nullptr
// grid (ceil(OT/128), n), 256 threads
__global__ void __launch_bounds__(256)
qadapter_pool_kernel(const float* __restrict__ map, float* __restrict__ out_f32,
                     op16_t* __restrict__ out_tm, long tm_stride_w, long tm_off, int H, int W,
                     int OF, int OT, int det, const QAdapterDev ad) {
  // per-CTA tables: the bin ranges (same float expressions as before, evaluated once per row / column instead of
  // once per output) and the <= 8 source columns this CTA's 128 output columns touch, staged in shared memory
  __shared__ int2 frange[128];          // [OF <= 128] source rows [fs, fe) of output row f
  __shared__ int2 trange[128];          // source columns [ts, te) of output column t0 + i
  __shared__ float cols[128 * 9];       // map[:, c0 : c0 + nc], row pitch 9
  const long n = blockIdx.y;
  const int t0 = blockIdx.x * 128;
  const int t1 = min(OT, t0 + 128);
  const int c0 = static_cast<int>(floorf(static_cast<float>(t0 * W) / OT));
  const int c1 = min(W, static_cast<int>(ceilf(static_cast<float>(t1 * W) / OT)));
  const int nc = c1 - c0;                               // <= 8 for W=128, OT=3000
  const float* src = map + n * static_cast<long>(H) * W;
  const float g = ad.gamma[det], be = ad.beta[det];
  float* out32 = out_f32 ? out_f32 + n * static_cast<long>(OF) * OT : nullptr;
  op16_t* outtm = out_tm ? out_tm + (n * tm_stride_w + tm_off) * static_cast<long>(OT + 2) * OF : nullptr;
  const bool staged = (OF <= 128 && H <= 128 && nc <= 9);
  if (staged) {
    for (int f = threadIdx.x; f < OF; f += 256)
      frange[f] = make_int2(static_cast<int>(floorf(static_cast<float>(f * H) / OF)),
                            static_cast<int>(ceilf(static_cast<float>((f + 1) * H) / OF)));
    for (int i = threadIdx.x; i < t1 - t0; i += 256) {
      const int t = t0 + i;
      trange[i] = make_int2(static_cast<int>(floorf(static_cast<float>(t * W) / OT)),
                            static_cast<int>(ceilf(static_cast<float>((t + 1) * W) / OT)));
    }
    for (int i = threadIdx.x; i < H * nc; i += 256) {
      const int y = i / nc, x = i - y * nc;
      cols[y * 9 + x] = __ldg(src + y * W + c0 + x);
    }
    __syncthreads();
  }
  // Time-major 16-bit output (the encoder's input), fast path: 3000 output columns come from only 128 source
  // columns, so consecutive output columns repeat -- an output column is determined by its source range
  // [ts, te) (one or two columns).  Each distinct column vector is computed once (same expression and summation
  // order as below) and the rows are then streamed out with 16-byte stores.
  __shared__ __align__(16) float vecs[16 * 128];        // [class][OF], class = (ts - c0) * 2 + (te - ts - 1)
  const bool fast_tm = staged && outtm != nullptr && out32 == nullptr && (OF % 8) == 0 && nc <= 8 && W <= OT;
  if (fast_tm) {
    for (int i = threadIdx.x; i < 2 * nc * OF; i += 256) {
      const int cls = i / OF, f = i - cls * OF;
      const int xs = c0 + (cls >> 1), xe = xs + 1 + (cls & 1);
      float v = 0.f;
      if (xe <= c1) {
        const int fs = frange[f].x, fe = frange[f].y;
        float sum = 0.f;
        for (int y = fs; y < fe; ++y)
          for (int x = xs; x < xe; ++x) sum += cols[y * 9 + (x - c0)];
        v = sum / static_cast<float>((fe - fs) * (xe - xs));
        v = ad.scale * v + ad.bias;
        v = v * g + be;
      }
      vecs[cls * 128 + f] = v;
    }
    __syncthreads();
    const int of8 = OF / 8;
    for (int idx = threadIdx.x; idx < of8 * (t1 - t0); idx += 256) {
      const int tt = idx / of8, f8 = idx - tt * of8;
      const int ts = trange[tt].x, te = trange[tt].y;
      const float* v = vecs + ((ts - c0) * 2 + (te - ts - 1)) * 128 + 8 * f8;
      uint4 pk;
      pk.x = pack_op16x2(v[0], v[1]); pk.y = pack_op16x2(v[2], v[3]);
      pk.z = pack_op16x2(v[4], v[5]); pk.w = pack_op16x2(v[6], v[7]);
      *reinterpret_cast<uint4*>(outtm + static_cast<long>(t0 + tt + 1) * OF + 8 * f8) = pk;
    }
  } else
  for (int idx = threadIdx.x; idx < OF * (t1 - t0); idx += 256) {
    const int tt = idx / OF, f = idx - tt * OF;          // f fastest: coalesced time-major stores
    const int t = t0 + tt;
    int fs, fe, ts, te;
    float sum = 0.f;
    if (staged) {
      fs = frange[f].x; fe = frange[f].y; ts = trange[tt].x; te = trange[tt].y;
      for (int y = fs; y < fe; ++y)
        for (int x = ts; x < te; ++x) sum += cols[y * 9 + (x - c0)];
    } else {
      fs = static_cast<int>(floorf(static_cast<float>(f * H) / OF));
      fe = static_cast<int>(ceilf(static_cast<float>((f + 1) * H) / OF));
      ts = static_cast<int>(floorf(static_cast<float>(t * W) / OT));
      te = static_cast<int>(ceilf(static_cast<float>((t + 1) * W) / OT));
      for (int y = fs; y < fe; ++y)
        for (int x = ts; x < te; ++x) sum += __ldg(src + y * W + x);
    }
    float v = sum / static_cast<float>((fe - fs) * (te - ts));
    v = ad.scale * v + ad.bias;
    v = v * g + be;
    if (out32) out32[static_cast<long>(f) * OT + t] = v;
    if (outtm) outtm[static_cast<long>(t + 1) * OF + f] = float_to_op16(v);
  }
  if (outtm && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < OF; i += 256) {
      outtm[i] = float_to_op16(0.f);
      outtm[static_cast<long>(OT + 1) * OF + i] = float_to_op16(0.f);
    }
  }
}

}  // namespace gww
