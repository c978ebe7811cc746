// Helpers shared by the tensor-core decoder kernels (forward and backward recurrences).
#pragma once
#include "decoder_common.cuh"
#include "tc_common.cuh"

namespace zeggs {

// byte offset of element (row, k) inside an image whose k-block tiles have `rows` rows
__host__ __device__ inline size_t img_off(int rows, int row, int k) {
  const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
  return (size_t)kb * rows * 128 + (size_t)row * 128 + (size_t)((c ^ (row & 7)) << 4) + (size_t)e * 2;
}


__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- thread-block clusters: rank / size, cluster barrier, multicast bulk copy (one L2 read feeds every CTA of the cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// the copy lands at the same shared-memory offset in every CTA of `mask` and completes `bytes` on each one's mbarrier
__device__ __forceinline__ void bulk_g2s_multicast(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[NC]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<4>(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}


// issue-only TMEM loads (no wait) and the 4-accumulator sum built on them: four loads in flight, ONE wait
template <int NC> __device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[NC]);
template <> __device__ __forceinline__ void tmem_ld_issue<8>(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tmem_ld_issue<4>(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
// v = acc0 + acc1 + acc2 + acc3 (accumulators `stride` columns apart), NC columns starting at taddr
template <int NC>
__device__ __forceinline__ void tmem_ld4_sum(uint32_t taddr, uint32_t stride, float (&v)[NC]) {
  uint32_t r0[NC], r1[NC], r2[NC], r3[NC];
  tmem_ld_issue<NC>(taddr, r0); tmem_ld_issue<NC>(taddr + stride, r1);
  tmem_ld_issue<NC>(taddr + 2 * stride, r2); tmem_ld_issue<NC>(taddr + 3 * stride, r3);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NC; ++i)
    v[i] = ((__uint_as_float(r0[i]) + __uint_as_float(r1[i])) + __uint_as_float(r2[i])) + __uint_as_float(r3[i]);
}

// same with NA (1, 2, 4 or 8) accumulators
template <int NC, int NA>
__device__ __forceinline__ void tmem_ldn_sum(uint32_t taddr, uint32_t stride, float (&v)[NC]) {
  if (NA == 1) {
    uint32_t r0[NC];
    tmem_ld_issue<NC>(taddr, r0);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r0[i]);
  } else if (NA == 2) {
    uint32_t r0[NC], r1[NC];
    tmem_ld_issue<NC>(taddr, r0); tmem_ld_issue<NC>(taddr + stride, r1);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r0[i]) + __uint_as_float(r1[i]);
  } else {
    tmem_ld4_sum<NC>(taddr, stride, v);
    if (NA == 8) {
      float u_[NC];
      tmem_ld4_sum<NC>(taddr + 4 * stride, stride, u_);
#pragma unroll
      for (int i = 0; i < NC; ++i) v[i] += u_[i];
    }
  }
}

// bf16-engine gate math: exp via the SFU (relative error ~1e-6, far below the bf16 operand rounding)
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

// grid barrier split in two halves: the epilogue warp arrives, the activation loader waits
__device__ __forceinline__ void grid_arrive(unsigned* counter) {
  __syncwarp();   // the lanes' stores happen-before lane 0's release (cumulative at gpu scope)
  if ((threadIdx.x & 31) == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void grid_wait(const unsigned* counter, unsigned target) {
  long long t0 = clock64();
  while (ld_acquire_u32(counter) < target) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// write U consecutive bf16 values (units j0..j0+U-1 of sample row b) into an activation image
template <int U>
__device__ __forceinline__ void store_img_units(uint8_t* img, int b, int j0, const float (&h)[U]) {
  __nv_bfloat16 t[U];
#pragma unroll
  for (int i = 0; i < U; ++i) t[i] = __float2bfloat16_rn(h[i]);
  uint8_t* p = img + img_off(32, b, j0);
  if (U == 8) *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(t);
  else *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(t);
}


}  // namespace zeggs
