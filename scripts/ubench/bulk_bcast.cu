// Micro-benchmark: how fast does a 64 KB activation image reach the shared memory of 128 CTAs that all read it right after a grid
// barrier (the all-to-all exchange of the recurrence kernels)?  Variables: freshly written by the 128 CTAs vs static, shared vs
// per-CTA private source, chunking (1 x 64 KB, 4 x 16 KB, 16 x 4 KB), one issuing thread vs four.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bcast bulk_bcast.cu ; run: ./bulk_bcast
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { for (uint32_t i = 0; i < 20000000u; ++i) if (mbar_try(b, parity)) return; __trap(); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory"); return v; }

constexpr int IMG = 65536;

// mode bits: 1 = write the image before the barrier (fresh), 2 = private source per CTA, chunks = number of bulk copies, nthr = issuing threads
__global__ void __launch_bounds__(128, 1) bcast_kernel(uint8_t* img, uint8_t* priv, unsigned* bar, long long* out, int reps, int fresh, int priv_src,
                                                       int chunks, int nthr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + IMG);          // [16]
  const int c = blockIdx.x, G = gridDim.x;
  if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  __syncthreads();
  const uint32_t cb = IMG / chunks;
  long long tot_first = 0, tot_last = 0;
  for (int r = 0; r < reps; ++r) {
    if (fresh && threadIdx.x < 32) {
      // this CTA's piece: 16 bytes in each of 32 different 128-byte lines (like 8 bf16 units of 32 sample rows), moving with r
      const int kb = (c * 8) >> 6, chunk = ((c * 8) & 63) >> 3, row = threadIdx.x;
      uint4 v = make_uint4(r, c, row, 7);
      *reinterpret_cast<uint4*>(img + (size_t)kb * 4096 + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(bar, 1u);
      const unsigned target = (unsigned)(r + 1) * G;
      long long t0 = clock64();
      while (ld_acq(bar) < target) { if (clock64() - t0 > 4000000000LL) __trap(); }
    }
    __syncthreads();
    asm volatile("fence.proxy.async;\n" ::: "memory");
    const uint8_t* src = priv_src ? priv + (size_t)c * IMG : img;
    const long long t0 = clock64();
    if ((int)threadIdx.x < nthr * 32 && (threadIdx.x & 31) == 0) {
      const int w = threadIdx.x >> 5;
      for (int k = w; k < chunks; k += nthr) {
        const int bi = k & 15;
        mbar_expect(&bars[bi], cb);
        bulk_g2s(smem + (size_t)k * cb, src + (size_t)k * cb, cb, &bars[bi]);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long tf = 0, tl = 0;
      const int nb = chunks < 16 ? chunks : 16;
      // chunk k signals barrier k & 15 (with 16 chunks every barrier gets exactly one); wait in order
      for (int k = 0; k < nb; ++k) {
        // with > 16 chunks a barrier would need several phases; this benchmark keeps chunks <= 16
        mbar_wait(&bars[k], (uint32_t)(r & 1));
        const long long t = clock64() - t0;
        if (k == 0) tf = t;
        tl = t;
      }
      tot_first += tf; tot_last += tl;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[c * 2] = tot_first / reps; out[c * 2 + 1] = tot_last / reps; }
}

int main() {
  int dev = 0; CK(cudaSetDevice(dev));
  uint8_t *img, *priv; unsigned* bar; long long* out;
  const int G = 128;
  CK(cudaMalloc(&img, IMG)); CK(cudaMalloc(&priv, (size_t)G * IMG)); CK(cudaMalloc(&bar, 4)); CK(cudaMalloc(&out, G * 2 * sizeof(long long)));
  CK(cudaMemset(img, 1, IMG)); CK(cudaMemset(priv, 2, (size_t)G * IMG));
  const size_t smem = IMG + 256 + 100 * 1024;       // + padding so that only one CTA fits per SM
  CK(cudaFuncSetAttribute(bcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  struct Cfg { const char* name; int fresh, priv, chunks, nthr; } cfgs[] = {
    {"shared, fresh,  4 x 16 KB, 1 thread ", 1, 0, 4, 1},  {"shared, static, 4 x 16 KB, 1 thread ", 0, 0, 4, 1},
    {"private,static, 4 x 16 KB, 1 thread ", 0, 1, 4, 1},  {"shared, fresh, 16 x  4 KB, 1 thread ", 1, 0, 16, 1},
    {"shared, fresh,  1 x 64 KB, 1 thread ", 1, 0, 1, 1},  {"shared, fresh,  4 x 16 KB, 4 threads", 1, 0, 4, 4},
    {"shared, fresh, 16 x  4 KB, 4 threads", 1, 0, 16, 4}, {"private,static,16 x  4 KB, 4 threads", 0, 1, 16, 4},
    {"shared, fresh,  8 x  8 KB, 4 threads", 1, 0, 8, 4},  {"shared, static,16 x  4 KB, 4 threads", 0, 0, 16, 4},
  };
  long long h[G * 2];
  for (auto& cf : cfgs) {
    CK(cudaMemset(bar, 0, 4));
    int reps = 200;
    void* args[] = {&img, &priv, &bar, &out, &reps, &cf.fresh, &cf.priv, &cf.chunks, &cf.nthr};
    CK(cudaLaunchCooperativeKernel((void*)bcast_kernel, dim3(G), dim3(128), args, smem, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
    long long f0 = h[0], l0 = h[1], fm = 0, lm = 0, ls = 0;
    for (int i = 0; i < G; ++i) { if (h[2 * i] > fm) fm = h[2 * i]; if (h[2 * i + 1] > lm) lm = h[2 * i + 1]; ls += h[2 * i + 1]; }
    printf("%s | CTA0 first chunk %6lld, all 64 KB %6lld cycles | mean over CTAs %6lld | slowest CTA %6lld (first %6lld) | %.1f B/clk/SM\n",
           cf.name, f0, l0, ls / G, lm, fm, (double)IMG / (double)(ls / G));
  }
  return 0;
}
