// Micro-benchmark: do tcgen05.ld (TMEM -> registers) and MUFU.EX2 overlap on one SM?
// 8 warps: warps 0-3 run exp2 chains, warps 4-7 stream TMEM loads; each role is timed alone and together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gw_whisper_b200/csrc -o tools/ubench/tmem_mufu.bin tools/ubench/tmem_mufu.cu
#include "ptx.cuh"
#include <cstdio>
using namespace gww;

// mode bit 0: MUFU warps active, bit 1: TMEM warps active
__global__ void __launch_bounds__(256) k(int mode, unsigned long long* cyc, float* sink, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(smem_u32(&tptr)); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c][i] = threadIdx.x + i + c;
  if (warp >= 4) {
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_st32(base + 32 * c, v[c]);
    tmem_wait_st();
  }
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
  if (warp < 4) {
    if (mode & 1) {
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = -0.001f * (threadIdx.x + i);
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {          // 128 exp2 per iteration (one S tile row)
#pragma unroll
        for (int rep = 0; rep < 8; ++rep)
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += f[i];
    }
  } else {
    if (mode & 2) {
#pragma unroll 1
      for (int it = 0; it < iters; ++it) {          // 128 columns per iteration (one S tile row)
        tmem_ld32(base + 0, v[0]); tmem_ld32(base + 32, v[1]); tmem_ld32(base + 64, v[2]); tmem_ld32(base + 96, v[3]);
        tmem_wait_ld();
        acc += __uint_as_float(v[0][0] ^ v[1][1] ^ v[2][2] ^ v[3][3]);
      }
    }
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  const long long t2 = clock64();
  if (threadIdx.x == 0) { cyc[2 * blockIdx.x] = (unsigned long long)(t2 - t0); }
  if (threadIdx.x == 128) { cyc[2 * blockIdx.x + 1] = (unsigned long long)(t1 - t0); }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}

int main() {
  unsigned long long* cyc; float* sink;
  const int nb = 148, iters = 2048;
  cudaMalloc(&cyc, nb * 16); cudaMalloc(&sink, nb * 256 * 4);
  for (int mode = 1; mode <= 3; ++mode) {
    k<<<nb, 256>>>(mode, cyc, sink, iters);
    k<<<nb, 256>>>(mode, cyc, sink, iters);
    cudaDeviceSynchronize();
    unsigned long long h[296];
    cudaMemcpy(h, cyc, nb * 16, cudaMemcpyDeviceToHost);
    double all = 0, tm = 0;
    for (int i = 0; i < nb; ++i) { all += (double)h[2 * i]; tm += (double)h[2 * i + 1]; }
    printf("mode %d (%s): CTA cycles/iter %.1f   TMEM-warp cycles/iter %.1f   err=%s\n", mode,
           mode == 1 ? "MUFU only, 4 warps x 128 ex2" : (mode == 2 ? "TMEM only, 4 warps x 128 cols" : "both concurrently"),
           all / nb / iters, tm / nb / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
