// (1) where do the 64 rows of an M=64 tcgen05.mma accumulator live in TMEM?  (2) issue/execute rate M=64 vs M=128, N=32
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/tc_common.cuh"
using namespace zeggs;
__device__ size_t img_off(int row, int k) { int kb = k >> 6, c = (k & 63) >> 3, e = k & 7; return (size_t)kb * 0 + (size_t)row * 128 + (size_t)((c ^ (row & 7)) << 4) + e * 2; }
__global__ void __launch_bounds__(128, 1) k(int M, int nmma, float* out, long long* tm, int N = 32) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  uint8_t* A = smem; uint8_t* B = smem + 64 * 1024;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) *reinterpret_cast<__nv_bfloat16*>(A + img_off(r, 0)) = __float2bfloat16_rn((float)(r + 1));
  for (int n = threadIdx.x; n < 32; n += blockDim.x) *reinterpret_cast<__nv_bfloat16*>(B + img_off(n, 0)) = __float2bfloat16_rn(1.0f);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async;\n" ::: "memory");
  tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_bf16_f32(M, N);
    const uint64_t da = make_smem_desc_sw128(A), db = make_smem_desc_sw128(B);
    long long t0 = clock64();
    for (int i = 0; i < nmma; i += 4) {
      if (elect_one_sync()) {
        umma_bf16(0, da, db, idesc, i > 0);
        umma_bf16(N, da + 2, db + 2, idesc, i > 0);
        umma_bf16(2 * N, da + 4, db + 4, idesc, i > 0);
        umma_bf16(3 * N, da + 6, db + 6, idesc, i > 0);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (elect_one_sync()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0) { tm[0] = t1 - t0; tm[1] = t2 - t0; }
  }
  __syncthreads();
  tc_fence_after_sync();
  uint32_t v[32];
  tmem_ld_32x32b_x32(((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x] = __uint_as_float(v[0]);
  tc_fence_before_sync(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after_sync(); tmem_dealloc(0, 512); }
}
int main() {
  float* d; long long* t; cudaMalloc(&d, 512); cudaMalloc(&t, 16); float h[128]; long long ht[2];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int M : {128, 64}) {
    k<<<1, 128, 200 * 1024>>>(M, 4, d, t);       // one round: D = A*B with only k=0 non-zero -> D[r][n] = r+1
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 512, cudaMemcpyDeviceToHost);
    printf("M=%d (%s): TMEM lane -> accumulated value (row+1):\n", M, cudaGetErrorString(e));
    for (int i = 0; i < 128; ++i) printf("%g%s", h[i], (i % 32 == 31) ? "\n" : " ");
    k<<<1, 128, 200 * 1024>>>(M, 4096, d, t);
    cudaDeviceSynchronize();
    cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
    printf("M=%d N=32: issue %.1f cyc/mma, complete %.1f cyc/mma\n", M, ht[0] / 4096.0, ht[1] / 4096.0);
    for (int N : {16, 48, 64, 128}) {
      k<<<1, 128, 200 * 1024>>>(M, 4096, d, t, N);
      cudaDeviceSynchronize();
      cudaMemcpy(ht, t, 16, cudaMemcpyDeviceToHost);
      printf("M=%d N=%d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", M, N, ht[0] / 4096.0, ht[1] / 4096.0);
    }
  }
  return 0;
}
