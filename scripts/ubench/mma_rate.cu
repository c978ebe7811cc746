// microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, SS) as a function of N and of the number of rotating accumulators
#include <cstdio>
#include <cuda_runtime.h>
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/tc_common.cuh"
using namespace zeggs;
__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int nmma, int kadv, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async;\n" ::: "memory");
    const uint32_t idesc = make_idesc_bf16_f32(128, N);
    long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const int kb = (i / 4) % 16, ks = i % 4;
      uint64_t da = make_smem_desc_sw128(smem + kb * kadv) + (uint64_t)(ks * 2);
      uint64_t db = make_smem_desc_sw128(smem + 96 * 1024 + (kb % 8) * 4096) + (uint64_t)(ks * 2);
      umma_bf16(tmem + (uint32_t)((i % nacc) * N), da, db, idesc, i >= nacc);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}
int main() {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int nmma = 1024;
  for (int kadv : {4096, 16384})
    for (int N : {16, 32, 64, 128, 256})
      for (int nacc : {1, 2, 4}) {
        if (nacc * N > 512) continue;
        k<<<1, 128, 200 * 1024>>>(N, nacc, nmma, kadv, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("A k-block stride %5d  N=%3d nacc=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (%s)\n", kadv, N, nacc, (double)h[0] / nmma, (double)h[1] / nmma, cudaGetErrorString(e));
      }
  return 0;
}
