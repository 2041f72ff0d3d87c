// Micro-benchmark of the per-SM issue rates that bound the softmax / GELU epilogues on sm_100a:
// MUFU (ex2, tanh) in f32 / f16x2 / bf16x2, FFMA, HFMA2, FMNMX, F2FP conversions.
// Reports lane-results per clock per SM (clock64 inside one 512-thread CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes.bin tools/ubench/pipes.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(512) k(float seed, unsigned long long* cyc, float* sink) {
  uint32_t r[CHAINS];
  float f[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    f[i] = seed + 0.001f * (threadIdx.x + i);
    __half2 h = __floats2half2_rn(f[i], f[i] * 0.5f);
    r[i] = *reinterpret_cast<uint32_t*>(&h);
    if (OP == 2 || OP == 5 || OP == 9) {
      __nv_bfloat162 b = __floats2bfloat162_rn(f[i], f[i] * 0.5f);
      r[i] = *reinterpret_cast<uint32_t*>(&b);
    }
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(r[i]));
      if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) % CHAINS] ), "f"(seed));
      if (OP == 7) asm volatile("fma.rn.f32 %0, %0, 0f3F7FF000, 0f3A000000;" : "+f"(f[i]));
      if (OP == 8) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));
      if (OP == 9) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));
      if (OP == 10) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) % CHAINS]));
      if (OP == 11) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 12) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, %1;}" : "+r"(r[i]) : "f"(f[i]));
      if (OP == 13) asm volatile("{.reg .f32 t; mov.b32 t, %0; cvt.rn.f16x2.f32 %0, t, %1;}" : "+r"(r[i]) : "f"(f[i]));
      if (OP == 14) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed));
      if (OP == 15) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));   // mixed: 1 MUFU + 3 FFMA
                      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i+1)%CHAINS]) : "f"(f[(i + 2) % CHAINS] ), "f"(seed));
                      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i+3)%CHAINS]) : "f"(f[(i + 4) % CHAINS] ), "f"(seed));
                      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[(i+5)%CHAINS]) : "f"(f[(i + 6) % CHAINS] ), "f"(seed)); }
      if (OP == 16) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));
      if (OP == 17) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) % CHAINS]));
      if (OP == 18) asm volatile("{.reg .b16 lo, hi; .reg .f32 a; mov.b32 {lo,hi}, %0; cvt.f32.f16 a, lo; mov.b32 %0, a;}" : "+r"(r[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc += f[i] + __uint_as_float(r[i]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP>
void run(const char* name, int results_per_instr, int nthreads) {
  unsigned long long* cyc;
  float* sink;
  int nb = 148;
  cudaMalloc(&cyc, nb * 8);
  cudaMalloc(&sink, nb * 512 * 4);
  k<OP><<<nb, nthreads>>>(0.5f, cyc, sink);
  k<OP><<<nb, nthreads>>>(0.5f, cyc, sink);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < nb; ++i) avg += (double)h[i];
  avg /= nb;
  const double instr = (double)ITERS * CHAINS * nthreads * (OP == 15 ? 4 : 1);
  printf("%-28s threads=%3d  cycles=%9.0f  lane-instr/clk/SM=%7.2f  results/clk/SM=%7.2f  err=%s\n", name, nthreads, avg,
         instr / avg, instr * results_per_instr / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  for (int nt : {128, 256, 512}) {
    run<0>("ex2.f32", 1, nt);
    run<1>("ex2.f16x2", 2, nt);
    run<2>("ex2.bf16x2", 2, nt);
    run<3>("tanh.f32", 1, nt);
    run<4>("tanh.f16x2", 2, nt);
    run<5>("tanh.bf16x2", 2, nt);
    run<6>("ffma 3reg", 1, nt);
    run<7>("ffma imm", 1, nt);
    run<8>("hfma2.f16x2", 2, nt);
    run<9>("hfma2.bf16x2", 2, nt);
    run<10>("fmnmx.f32", 1, nt);
    run<11>("fadd.f32", 1, nt);
    run<14>("fmul.f32", 1, nt);
    run<12>("cvt.bf16x2.f32", 2, nt);
    run<13>("cvt.f16x2.f32", 2, nt);
    run<15>("mix ex2+3ffma", 1, nt);
    run<16>("hadd2.f16x2", 2, nt);
    run<17>("hmnmx2.bf16x2", 2, nt);
    run<18>("cvt.f32.f16", 1, nt);
  }
  return 0;
}
