// Micro-benchmark: cost of one tcgen05.mma (kind::f16, bf16 in, f32 accumulate, cta_group::1, M = 128)
// as a function of N and of where A comes from (shared memory or TMEM), measured as a long chain of
// accumulating MMAs on one accumulator, timed from first issue to the commit's mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I gw_whisper_b200/csrc -o tools/ubench/umma_shapes.bin tools/ubench/umma_shapes.cu
#include "ptx.cuh"
#include <cstdio>
using namespace gww;

template <int N, int TS, int BMN>
__global__ void __launch_bounds__(128) k(unsigned long long* cyc, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<512>(smem_u32(&tptr)); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, BMN);
    const uint64_t adesc = make_sw128_desc(smem_u32(smem));
    const uint64_t bdesc = make_sw128_desc(smem_u32(smem) + 16384);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (TS) umma_ts(tb, tb + 256 + 8 * kk, bdesc + (BMN ? 128 * kk : 2 * kk), idesc, 1u);
            else umma_ss(tb, adesc + 2 * kk, bdesc + (BMN ? 128 * kk : 2 * kk), idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), rep & 1);
      t1 = clock64();
    }
    if (threadIdx.x == 32) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

template <int N, int TS, int BMN>
void run(const char* name) {
  unsigned long long* cyc;
  const int nb = 148, iters = 512;
  cudaMalloc(&cyc, nb * 8);
  cudaFuncSetAttribute(k<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<N, TS, BMN><<<nb, 128, 65536>>>(cyc, iters);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, nb * 8, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < nb; ++i) avg += (double)h[i]; avg /= nb;
  const double per = avg / (iters * 4);
  printf("%-36s N=%3d  cycles/MMA=%7.1f  MAC/clk/SM=%7.0f  err=%s\n", name, N, per, 128.0 * N * 16 / per,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(cyc);
}

int main() {
  run<256, 0, 0>("SS, B K-major");
  run<128, 0, 0>("SS, B K-major");
  run<64, 0, 0>("SS, B K-major");
  run<32, 0, 0>("SS, B K-major");
  run<128, 1, 0>("TS (A in TMEM), B K-major");
  run<64, 1, 0>("TS (A in TMEM), B K-major");
  run<64, 0, 1>("SS, B MN-major (the V operand)");
  run<64, 1, 1>("TS, B MN-major (the P.V MMA)");
  run<128, 1, 1>("TS, B MN-major");
  return 0;
}
