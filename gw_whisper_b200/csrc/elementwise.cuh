// Bandwidth-bound helpers of the encoder path: LayerNorm (HF modeling_whisper.py:393,403,643),
// the pooled MLP classifier heads (MLGWSC-1/inference.py:371-382, Signal_vs_Noise/src/model.py:9-20,
// 35-47, Glitch_classification/src/model.py:10-21) and the threshold -> trigger compaction that
// replaces the per-element python loop at MLGWSC-1/inference.py:484-487.
#pragma once
#include "ptx.cuh"

namespace gww {

// ---------------------------------------------------------------------------------------------
// LayerNorm over the last dim (d = 128*NV), one warp per row, two-pass in registers.
// in  : f32, row i is read at in + (in_off + i*in_stride) * d
// out : bf16 or f32, dense [rows, d]
template <int NV, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, OutT* __restrict__ out,
                 const float* __restrict__ gamma, const float* __restrict__ beta, long rows,
                 long in_off, long in_stride, float eps) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(in + (in_off + row * in_stride) * D);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = src[lane + 32 * i];
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    const float o0 = fmaf(v[i].x * rstd, g.x, b.x), o1 = fmaf(v[i].y * rstd, g.y, b.y);
    const float o2 = fmaf(v[i].z * rstd, g.z, b.z), o3 = fmaf(v[i].w * rstd, g.w, b.w);
    if constexpr (sizeof(OutT) == 2) {
      uint2 pk = make_uint2(pack_op16x2(o0, o1), pack_op16x2(o2, o3));
      reinterpret_cast<uint2*>(out + row * D)[lane + 32 * i] = pk;
    } else {
      reinterpret_cast<float4*>(out + row * D)[lane + 32 * i] = make_float4(o0, o1, o2, o3);
    }
  }
}

// mean over tokens (use_last_token=False branch, MLGWSC-1/inference.py:390): hs [Bt, T, d] f32
template <int DUMMY = 0>
__global__ void __launch_bounds__(256)
mean_pool_kernel(const float* __restrict__ hs, float* __restrict__ out, int T, int d) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += hs[(static_cast<size_t>(b) * T + t) * d + c];
    out[static_cast<size_t>(b) * d + c] = acc / static_cast<float>(T);
  }
}

// ---------------------------------------------------------------------------------------------
// Last-token-only tail of the final encoder layer.  The reference consumes only
// last_hidden_state[:, -1, :] (MLGWSC-1/inference.py:390, Signal_vs_Noise/src/model.py:25-26); in the
// final layer that value depends on every token's K and V but only on the LAST token's query,
// attention row, out-projection, MLP and final LayerNorm.  Same arithmetic as HF
// modeling_whisper.py:215-238 for that one row, in fp32 on CUDA cores (1500 x 64 MACs per head).
//   qkv [nc, T, 3d] bf16 (q pre-scaled) -> out [nc, d] bf16 ; grid (heads, nc), 256 threads
__global__ void __launch_bounds__(256)
last_row_attention_kernel(const op16_t* __restrict__ qkv, op16_t* __restrict__ out, int T,
                          int d) {
  extern __shared__ __align__(16) float lra_smem[];   // [T] scores | [8][64] partial outputs | [16] reductions
  float* sc = lra_smem;
  float* part = sc + ((T + 3) & ~3);
  float* red = part + 8 * 64;
  const int h = blockIdx.x;
  const long b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const op16_t* head0 = qkv + b * static_cast<long>(T) * 3 * d + h * 64;          // q of token 0, this head
  const op16_t* base = head0 + 2 * lane;
  const op16_t* vb = base + 2 * d;
  // scores: one thread per key (its 128-byte K row against the query row held in registers); r1 spread each key over the
  // 32 lanes and paid five shuffle rounds per dot product
  float q[64];
  {
    const uint4* qr = reinterpret_cast<const uint4*>(head0 + static_cast<long>(T - 1) * 3 * d);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = qr[c];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = op16x2_to_float2(*reinterpret_cast<const op16x2_t*>(&w[e]));
        q[8 * c + 2 * e] = f.x;
        q[8 * c + 2 * e + 1] = f.y;
      }
    }
  }
  float wmax = -INFINITY;
  for (int k = threadIdx.x; k < T; k += 256) {
    const uint4* kr = reinterpret_cast<const uint4*>(head0 + d + static_cast<long>(k) * 3 * d);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 v = kr[c];
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = op16x2_to_float2(*reinterpret_cast<const op16x2_t*>(&w[e]));
        d0 = fmaf(q[8 * c + 2 * e], f.x, d0);
        d1 = fmaf(q[8 * c + 2 * e + 1], f.y, d1);
      }
    }
    const float dot = d0 + d1;
    sc[k] = dot;
    wmax = fmaxf(wmax, dot);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) red[warp] = wmax;
  __syncthreads();
  float m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  float lsum = 0.f;
  for (int k = threadIdx.x; k < T; k += 256) {
    const float pk = __expf(sc[k] - m);
    sc[k] = pk;
    lsum += pk;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[8 + warp] = lsum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[8 + i];
  float2 acc = make_float2(0.f, 0.f);
  for (int k0 = warp * 4; k0 < T; k0 += 32) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = min(k0 + u, T - 1);
      const float pk = (k0 + u < T) ? sc[k] : 0.f;
      const float2 vv = op16x2_to_float2(*reinterpret_cast<const op16x2_t*>(vb + static_cast<long>(k) * 3 * d));
      acc.x = fmaf(pk, vv.x, acc.x);
      acc.y = fmaf(pk, vv.y, acc.y);
    }
  }
  part[warp * 64 + 2 * lane] = acc.x;
  part[warp * 64 + 2 * lane + 1] = acc.y;
  __syncthreads();
  if (threadIdx.x < 64) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += part[i * 64 + threadIdx.x];
    out[b * d + h * 64 + threadIdx.x] = float_to_op16(v / tot);
  }
}

// x_last[i, :] = x[i*T + T-1, :]   (f32 residual rows of the last token)
__global__ void __launch_bounds__(128)
gather_last_rows_kernel(const float* __restrict__ x, float* __restrict__ x_last, int T, int d) {
  const long i = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(x + (i * T + (T - 1)) * d);
  float4* dst = reinterpret_cast<float4*>(x_last + i * d);
  for (int c = threadIdx.x; c < d / 4; c += 128) dst[c] = src[c];
}

// ---------------------------------------------------------------------------------------------
// Pooled classifier head: x [B, in0] f32 -> Linear(+ReLU) x (L-1) -> Linear -> optional softmax.
// One launch per Linear so every layer fills the machine: CTA = 8 output neurons x 16 windows,
// a warp owns one neuron, its lanes stride over the input with float4 loads (coalesced weight rows,
// conflict-free smem reads of the 16 staged input rows), shuffle reduction, fp32 throughout.
constexpr int kHeadMaxLayers = 6;
constexpr int kHeadWin = 16;
constexpr int kHeadMaxWidth = 1536;
struct HeadParams {
  int n_layers;
  int dims[kHeadMaxLayers + 1];       // dims[0]=in, dims[i+1]=out of layer i
  const float* w[kHeadMaxLayers];     // [out, in] row-major (nn.Linear)
  const float* b[kHeadMaxLayers];
  int softmax;                        // apply softmax over the last layer's outputs
  int B;
};

__global__ void __launch_bounds__(256)
head_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                   const float* __restrict__ bias, float* __restrict__ y, int B, int din, int dout,
                   int relu) {
  extern __shared__ __align__(16) float xs[];     // [kHeadWin][din]
  const int b0 = blockIdx.y * kHeadWin;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kHeadWin * din; i += 256) {
    const int q = i / din, c = i - q * din;
    xs[i] = (b0 + q < B) ? x[static_cast<size_t>(b0 + q) * din + c] : 0.f;
  }
  __syncthreads();
  const int j = blockIdx.x * 8 + warp;
  if (j >= dout) return;
  float acc[kHeadWin];
#pragma unroll
  for (int q = 0; q < kHeadWin; ++q) acc[q] = 0.f;
  const float* wr = W + static_cast<size_t>(j) * din;
  if ((din & 3) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(wr);
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    const int n4 = din >> 2;
    for (int i = lane; i < n4; i += 32) {
      const float4 wv = __ldg(w4 + i);
#pragma unroll
      for (int q = 0; q < kHeadWin; ++q) {
        const float4 xv = x4[q * n4 + i];
        acc[q] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[q]))));
      }
    }
  } else {
    for (int i = lane; i < din; i += 32) {
      const float wv = __ldg(wr + i);
#pragma unroll
      for (int q = 0; q < kHeadWin; ++q) acc[q] = fmaf(wv, xs[q * din + i], acc[q]);
    }
  }
  float mine = 0.f;
#pragma unroll
  for (int q = 0; q < kHeadWin; ++q) {
    float v = acc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == q) mine = v;
  }
  if (lane < kHeadWin && b0 + lane < B) {
    float v = mine + bias[j];
    if (relu) v = fmaxf(v, 0.f);
    y[static_cast<size_t>(b0 + lane) * dout + j] = v;
  }
}

// nn.Softmax(dim=1) over C (<= 64) outputs, in place or to `out`
__global__ void __launch_bounds__(256)
row_softmax_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C) {
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= B) return;
  const float* p = in + static_cast<size_t>(r) * C;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, p[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(p[c] - m);
  for (int c = 0; c < C; ++c) out[static_cast<size_t>(r) * C + c] = expf(p[c] - m) / s;
}

// ---------------------------------------------------------------------------------------------
// Ordered stream compaction of windows whose score (column 0 of out[B, C]) exceeds the threshold
// (strictly greater, MLGWSC-1/inference.py:484). Single CTA, ballot scan: output order == window
// order, so the host-side clustering sees the same sequence the reference's python loop builds.
__global__ void __launch_bounds__(1024)
threshold_compact_kernel(const float* __restrict__ out, int C, int n, float thr, long idx_base,
                         long* __restrict__ trig_idx, float* __restrict__ trig_score,
                         int* __restrict__ count, int capacity) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) base_s = *count;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int i = start + threadIdx.x;
    const float sc = (i < n) ? out[static_cast<size_t>(i) * C] : 0.f;
    const bool keep = (i < n) && (sc > thr);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    off += __popc(bal & ((1u << lane) - 1u));
    if (keep && off < capacity) {
      trig_idx[off] = idx_base + i;
      trig_score[off] = sc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_tot[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base_s;
}

}  // namespace gww
