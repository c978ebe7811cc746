// What paces a chain of small tcgen05 MMAs?  Variants: MMAs per commit, marching operand addresses, mbarrier wait before each group.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/tc_common.cuh"
using namespace zeggs;
// mode bits: 1 = commit after each group, 2 = march A/B addresses, 4 = try_wait on an (already complete) mbarrier before each group
__global__ void __launch_bounds__(128, 1) k(int M, int N, int group, int ngroups, int mode, long long* tm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2, done; __shared__ uint32_t slot;
  uint8_t* A = smem; uint8_t* B = smem + 128 * 1024;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;\n" ::: "memory");
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) mbar_arrive(&bar2);     // phase 0 of bar2 complete: waits on parity 0 pass immediately
    __syncwarp();
    const uint32_t idesc = make_idesc_bf16_f32(M, N);
    const uint64_t da0 = make_smem_desc_sw128(A), db0 = make_smem_desc_sw128(B);
    long long t0 = clock64();
    for (int g = 0; g < ngroups; ++g) {
      if (mode & 4) { mbar_wait(&bar2, 0); tc_fence_after_sync(); }
      const uint64_t da = da0 + ((mode & 2) ? (uint64_t)((g % 8) * 512) : 0), db = db0 + ((mode & 2) ? (uint64_t)((g % 8) * 512) : 0);
      if (elect_one_sync()) {
        for (int i = 0; i < group; ++i) {
          const uint64_t ka = (uint64_t)((i & 3) * 2 + (i >> 2) * 256), kb = (uint64_t)((i & 3) * 2 + (i >> 2) * (N * 8));
          umma_bf16((uint32_t)((i & 3) * N), da + ((mode & 2) ? ka : (uint64_t)((i & 3) * 2)), db + ((mode & 2) ? kb : (uint64_t)((i & 3) * 2)), idesc, (g | (i >> 2)) != 0);
        }
        if (mode & 1) umma_commit(&bar);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one_sync()) umma_commit(&done);
    __syncwarp();
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (lane == 0) { tm[0] = t1 - t0; tm[1] = t2 - t0; }
  }
  tc_fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(0, 512); }
}
int main() {
  long long* t; cudaMalloc(&t, 16); long long ht[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int NG = 512;
  for (int M : {64, 128}) for (int N : {24, 32}) for (int group : {4, 8, 16}) for (int mode : {0, 1, 2, 3, 7}) {
    k<<<1, 128, 200 * 1024>>>(M, N, group, NG, mode, t);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%2d group=%2d mode=%d (%s): issue %.1f cyc/mma, complete %.1f cyc/mma, %.0f cyc/group\n", M, N, group, mode,
           cudaGetErrorString(e), ht[0] / (double)(NG * group), ht[1] / (double)(NG * group), ht[1] / (double)NG);
  }
  return 0;
}
