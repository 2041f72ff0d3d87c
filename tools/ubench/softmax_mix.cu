// Micro-benchmark: the per-element instruction mix of the attention softmax (per PAIR of elements:
// 2 FFMA, 2 MUFU.EX2, 2 FADD, 1 F2FP, 1 FMNMX3) with no TMEM, no barriers, no tensor core.
// Question: can this mix keep the MUFU at 16 results/clk/SM with 1/2/4 warps per SM sub-partition?
// Variants drop instruction classes to see which pipe the mix is bound by.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/softmax_mix.bin tools/ubench/softmax_mix.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <stdint.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }

// MODE bits: 1 = FFMA scale, 2 = MUFU, 4 = FADD row sum, 8 = F2FP pack, 16 = FMNMX3 max
template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned long long* cyc, float* sink, int iters, float scale, float mneg) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = 0.001f * (threadIdx.x + i);
  float l0 = 0.f, l1 = 0.f, mx = -1e30f;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE & 16) {
#pragma unroll
      for (int i = 0; i < 64; i += 2) mx = fmaxf(mx, fmaxf(s[i], s[i + 1]));
    }
    const float mn = (MODE & 16) ? mneg + mx * 1e-30f : mneg;
#pragma unroll
    for (int i = 0; i < 64; i += 2) {
      float x0 = s[i], x1 = s[i + 1];
      if (MODE & 1) { x0 = fmaf(x0, scale, mn); x1 = fmaf(x1, scale, mn); }
      float p0 = x0, p1 = x1;
      if (MODE & 2) { p0 = ex2(x0); p1 = ex2(x1); }
      if (MODE & 4) { l0 += p0; l1 += p1; }
      if (MODE & 8) acc ^= pack(p0, p1);
      else acc ^= __float_as_uint(p0) ^ __float_as_uint(p1);
      s[i] = p0 * 0.5f + 0.25f;      // keep the chain data-dependent across iterations (1 FFMA per element)
      s[i + 1] = p1 * 0.5f + 0.25f;
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = l0 + l1 + mx + __uint_as_float(acc);
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int MODE>
void run(const char* name, int nthreads) {
  unsigned long long* cyc; float* sink;
  const int nb = 148, iters = 2048;
  cudaMalloc(&cyc, nb * 8); cudaMalloc(&sink, nb * 512 * 4);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters, 1.4426950408889634f, -3.0f);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters, 1.4426950408889634f, -3.0f);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += (double)h[i]; avg /= nb;
  const double elems = (double)iters * 64 * nthreads;
  printf("%-44s threads=%3d  cycles/64elem/warp-set=%8.1f  elements/clk/SM=%6.2f  err=%s\n", name, nthreads, avg / iters,
         elems / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int nt : {128, 256, 512}) {
    run<2>("MUFU only (+1 FFMA chain)", nt);
    run<3>("FFMA + MUFU", nt);
    run<7>("FFMA + MUFU + FADD", nt);
    run<15>("FFMA + MUFU + FADD + F2FP", nt);
    run<31>("FFMA + MUFU + FADD + F2FP + FMNMX (full)", nt);
    run<29>("full without MUFU", nt);
    run<27>("full without FADD", nt);
  }
  return 0;
}
