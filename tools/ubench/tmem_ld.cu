// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (lane-quarter = warp % 4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gw_whisper_b200/csrc -o tools/ubench/tmem_ld.bin tools/ubench/tmem_ld.cu
#include "ptx.cuh"
#include <cstdio>
using namespace gww;

template <int MODE>
__global__ void __launch_bounds__(256) k(unsigned long long* cyc, float* sink, int iters) {
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
  // initialise the columns we read
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_st32(base + 32 * c, v[c]);
  tmem_wait_st();
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 128 columns loaded, one wait (the attention pattern)
      tmem_ld32(base + 0, v[0]); tmem_ld32(base + 32, v[1]); tmem_ld32(base + 64, v[2]); tmem_ld32(base + 96, v[3]);
      tmem_wait_ld();
    } else if (MODE == 1) {   // 32 columns + wait each time (latency-exposed)
      tmem_ld32(base + 32 * (it & 3), v[0]);
      tmem_wait_ld();
    } else {                  // store 64 columns (P write pattern)
      tmem_st32(base + 256, v[0]); tmem_st32(base + 288, v[1]);
      tmem_wait_st();
    }
    acc += __uint_as_float(v[0][it & 31]);
  }
  const long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}

template <int MODE>
void run(const char* name, int nthreads, double bytes_per_iter_per_warp) {
  unsigned long long* cyc; float* sink;
  const int nb = 148, iters = 4096;
  cudaMalloc(&cyc, nb * 8); cudaMalloc(&sink, nb * 256 * 4);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += (double)h[i]; avg /= nb;
  const double bytes = bytes_per_iter_per_warp * iters * (nthreads / 32);
  printf("%-34s warps=%d cycles/iter=%8.1f  bytes/clk/SM=%8.1f  err=%s\n", name, nthreads / 32, avg / iters, bytes / avg,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int nt : {32, 128, 256}) {
    run<0>("ld 4x(32x32b.x32) + wait", nt, 32.0 * 128 * 4);
    run<1>("ld 1x(32x32b.x32) + wait", nt, 32.0 * 32 * 4);
    run<2>("st 2x(32x32b.x32) + wait", nt, 32.0 * 64 * 4);
  }
  return 0;
}
