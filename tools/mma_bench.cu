// mma_bench.cu — microbenchmark: cycles per tcgen05.mma (M=128, K=16, kind::f16) as a function of
// N, operand source (A from smem = SS, A from TMEM = TS) and B major-ness.  One CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o tools/mma_bench tools/mma_bench.cu
#include <cstdio>
#include "../nans-clip_b200/csrc/common.cuh"
using namespace nans;

__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tptr;
  if (warp == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 64 * 1024);
    // mode 0: SS K-major B ; 1: TS K-major B ; 2: SS MN-major B ; 3: TS MN-major B ; 4: SS, 8 distinct A/B chunks
    const bool ts = (mode & 1) != 0 && mode < 4;
    const bool mn = mode == 2 || mode == 3;
    const uint32_t idesc = make_idesc(0, 0, 0, mn ? 1 : 0, 128, N);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t off = (mode == 4) ? ((it & 7) * 16384 + k * 32) : k * 32;
          const uint64_t ad = make_smem_desc(a_addr + off, 16, 1024);
          const uint64_t bd = mn ? make_smem_desc(b_addr + k * 2048, 8192, 1024) : make_smem_desc(b_addr + off, 16, 1024);
          if (ts) mma_ts(tb, tb + 256 + k * 8, bd, idesc, 1u);
          else mma_ss(tb, ad, bd, idesc, 1u);
        }
      }
      tc_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"SS Kmajor", "TS Kmajor", "SS MNmajorB", "TS MNmajorB", "SS 8 chunks"};
  for (int mode = 0; mode < 5; ++mode)
    for (int N : {16, 32, 64, 128, 256}) {
      for (int grid : {1, 148}) {
        const int iters = 2000;
        bench<<<grid, 128, 200 * 1024>>>(N, mode, iters, d);
        bench<<<grid, 128, 200 * 1024>>>(N, mode, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("%-12s N=%3d grid=%3d : %7.1f clk/MMA  (%s)\n", names[mode], N, grid, double(h) / (iters * 4), cudaGetErrorString(e));
      }
    }
  return 0;
}
