// Replica of the recurrence kernels' MMA issue loop (ring slot of 2 k-blocks = 8 MMAs, M=64) with every mbarrier already
// complete: what does the loop cost per slot?  variant 0: as the kernel; 1: no waits; 2: one lane polls; 3: 16 MMAs per wait;
// 4: waits only (no MMAs); 5: no commit
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/tc_common.cuh"
using namespace zeggs;
template <int V> __device__ __forceinline__ void wait1(uint64_t* bar, uint32_t parity) {
  if (V >= 6) {   // non-blocking test first
    uint32_t done;
    asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  mbar_wait(bar, parity);
}
template <int V>
__global__ void __launch_bounds__(128, 1) k(int N, int nslots, long long* tm) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[8], empty[8], xf[4], done; __shared__ uint32_t slot;
  uint8_t* X = smem; uint8_t* ring = smem + 128 * 1024;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } for (int i = 0; i < 4; ++i) mbar_init(&xf[i], 1); mbar_init(&done, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;\n" ::: "memory");
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(64, N);
    const uint64_t dX = make_smem_desc_sw128(X), dRing = make_smem_desc_sw128(ring);
    const uint64_t bstep = (uint64_t)(N * 8);
    const uint32_t d0 = 256;
    const int SLOT = 2 * 40 * 128;
    long long t0 = clock64();
    uint32_t it = 0;
    for (int rep = 0; rep < nslots / 8; ++rep) {
      for (int kb = 0; kb < 16; kb += 2, ++it) {
        const uint32_t s = it % 6;
        if (V == 0 || V == 4 || V == 5 || V == 6 || V == 8) {
          if ((kb & 3) == 0) wait1<V>(&xf[kb >> 2], 1);
          wait1<V>(&full[s], 1);
          if (V != 8) tc_fence_after_sync();
        } else if (V == 2) {
          if (lane == 0) { if ((kb & 3) == 0) wait1<V>(&xf[kb >> 2], 1); wait1<V>(&full[s], 1); }
          __syncwarp();
          tc_fence_after_sync();
        } else if (V == 3 || V == 7) {
          if ((kb & 3) == 0) { wait1<V>(&xf[kb >> 2], 1); wait1<V>(&full[s], 1); tc_fence_after_sync(); }
        }
        const uint64_t da = dX + (uint64_t)kb * 256, db = dRing + (uint64_t)s * (SLOT >> 4);
        const bool acc0 = (kb | rep) > 0;
        if (elect_one_sync()) {
          if (V != 4) {
            umma_bf16(d0 + 0 * N, da + 0, db + 0, idesc, acc0);
            umma_bf16(d0 + 1 * N, da + 2, db + 2, idesc, acc0);
            umma_bf16(d0 + 2 * N, da + 4, db + 4, idesc, acc0);
            umma_bf16(d0 + 3 * N, da + 6, db + 6, idesc, acc0);
            umma_bf16(d0 + 0 * N, da + 256 + 0, db + bstep + 0, idesc, true);
            umma_bf16(d0 + 1 * N, da + 256 + 2, db + bstep + 2, idesc, true);
            umma_bf16(d0 + 2 * N, da + 256 + 4, db + bstep + 4, idesc, true);
            umma_bf16(d0 + 3 * N, da + 256 + 6, db + bstep + 6, idesc, true);
          }
          if (V != 5) umma_commit(&empty[s]);
        }
        __syncwarp();
      }
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
template <int V> void run(int N, long long* t) {
  long long ht[2]; const int NS = 1024;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  k<V><<<1, 128, 210 * 1024>>>(N, NS, t);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
  printf("variant %d N=%d (%s): issue %.0f cyc/slot, complete %.0f cyc/slot (8 MMAs)\n", V, N, cudaGetErrorString(e), ht[0] / (double)NS, ht[1] / (double)NS);
}
int main() {
  long long* t; cudaMalloc(&t, 16);
  for (int N : {24}) { run<0>(N, t); run<1>(N, t); run<2>(N, t); run<3>(N, t); run<4>(N, t); run<5>(N, t); run<6>(N, t); run<7>(N, t); run<8>(N, t); }
  return 0;
}
