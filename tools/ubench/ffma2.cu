// Micro-benchmark: fp32 FMA throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma2.bin tools/ubench/ffma2.cu
#include <cstdio>
#include <stdint.h>
template <int PACKED>
__global__ void __launch_bounds__(1024) k(unsigned long long* cyc, float* sink, int iters, float b, float c) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  uint64_t bb, cc;
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        uint64_t v, d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a[i]), "f"(a[i + 1]));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(v), "l"(bb), "l"(cc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
  }
  const long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}
template <int PACKED>
void run(const char* name, int nt) {
  unsigned long long* cyc; float* sink;
  const int nb = 148, iters = 4096;
  cudaMalloc(&cyc, nb * 8); cudaMalloc(&sink, nb * 1024 * 4);
  k<PACKED><<<nb, nt>>>(cyc, sink, iters, 0.999f, 0.001f);
  k<PACKED><<<nb, nt>>>(cyc, sink, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += (double)h[i]; avg /= nb;
  printf("%-10s threads=%4d  FMA/clk/SM=%7.1f  err=%s\n", name, nt, (double)iters * 16 * nt / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc); cudaFree(sink);
}
int main() {
  for (int nt : {256, 512, 1024}) { run<0>("FFMA", nt); run<1>("FFMA2", nt); }
  return 0;
}
