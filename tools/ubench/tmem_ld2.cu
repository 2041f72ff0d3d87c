// Micro-benchmark 2: tcgen05.ld cost per warp as a function of load size and of the number of warps,
// with the loaded values consumed through static register indices (tmem_ld.cu indexed the array
// dynamically, which put it in local memory and polluted its numbers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gw_whisper_b200/csrc -o tools/ubench/tmem_ld2.bin tools/ubench/tmem_ld2.cu
#include "ptx.cuh"
#include <cstdio>
using namespace gww;

// MODE 0: 1 x ld32 + wait   1: 2 x ld32 + wait   2: 4 x ld32 + wait   3: 4 x (ld32 + wait)
// MODE 4: 2 x ld16 + wait   5: 4 x ld32, wait, then 128 dependent FADDs (consume everything)
// MODE 6: st 2 x 32 + wait  7: ld 4x32 + wait + st 2x32 + wait (softmax skeleton)
template <int MODE>
__global__ void __launch_bounds__(512) k(unsigned long long* cyc, float* sink, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc<512>(smem_u32(&tptr)); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
  uint32_t v[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c][i] = __float_as_uint(1.0f + threadIdx.x + i + c);
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_st32(base + 32 * c, v[c]);
  tmem_wait_st();
  __syncthreads();
  const long long t0 = clock64();
  float acc = 0.f;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_ld32(base, v[0]); tmem_wait_ld();
      acc += __uint_as_float(v[0][0]) + __uint_as_float(v[0][31]);
    } else if (MODE == 1) {
      tmem_ld32(base, v[0]); tmem_ld32(base + 32, v[1]); tmem_wait_ld();
      acc += __uint_as_float(v[0][0]) + __uint_as_float(v[1][31]);
    } else if (MODE == 2) {
      tmem_ld32(base, v[0]); tmem_ld32(base + 32, v[1]); tmem_ld32(base + 64, v[2]); tmem_ld32(base + 96, v[3]);
      tmem_wait_ld();
      acc += __uint_as_float(v[0][0]) + __uint_as_float(v[1][31]) + __uint_as_float(v[2][7]) + __uint_as_float(v[3][9]);
    } else if (MODE == 3) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { tmem_ld32(base + 32 * c, v[c]); tmem_wait_ld(); acc += __uint_as_float(v[c][c]); }
    } else if (MODE == 4) {
      uint32_t a[16], b[16];
      tmem_ld16(base, a); tmem_ld16(base + 16, b); tmem_wait_ld();
      acc += __uint_as_float(a[0]) + __uint_as_float(b[15]);
    } else if (MODE == 5) {
      tmem_ld32(base, v[0]); tmem_ld32(base + 32, v[1]); tmem_ld32(base + 64, v[2]); tmem_ld32(base + 96, v[3]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[c][i]);
    } else if (MODE == 6) {
      tmem_st32(base + 256, v[0]); tmem_st32(base + 288, v[1]); tmem_wait_st();
      acc += 1.0f;
    } else {
      tmem_ld32(base, v[0]); tmem_ld32(base + 32, v[1]); tmem_ld32(base + 64, v[2]); tmem_ld32(base + 96, v[3]);
      tmem_wait_ld();
      acc += __uint_as_float(v[0][0]) + __uint_as_float(v[3][9]);
      tmem_st32(base + 256 - ((warp >> 2) & 1) * 64, v[1]); tmem_st32(base + 288 - ((warp >> 2) & 1) * 64, v[2]); tmem_wait_st();
    }
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
  const int nb = 148, iters = 2048;
  cudaMalloc(&cyc, nb * 8); cudaMalloc(&sink, nb * 512 * 4);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters);
  k<MODE><<<nb, nthreads>>>(cyc, sink, iters);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += (double)h[i]; avg /= nb;
  const double bytes = bytes_per_iter_per_warp * iters * (nthreads / 32);
  printf("%-40s warps=%2d cycles/iter=%8.1f  bytes/clk/SM=%8.1f  err=%s\n", name, nthreads / 32, avg / iters, bytes / avg,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int nt : {32, 128, 256, 512}) {
    run<0>("1 x ld32 + wait", nt, 4096.0);
    run<1>("2 x ld32 + wait", nt, 8192.0);
    run<2>("4 x ld32 + wait", nt, 16384.0);
    run<3>("4 x (ld32 + wait)", nt, 16384.0);
    run<4>("2 x ld16 + wait", nt, 4096.0);
    run<5>("4 x ld32 + wait + 128 FADD", nt, 16384.0);
    run<6>("st 2 x 32 + wait", nt, 8192.0);
    run<7>("ld 4x32 + wait + st 2x32 + wait", nt, 16384.0 + 8192.0);
  }
  return 0;
}
