// Generic fp32 SIMT GEMM for the batched (non-recurrent) linear layers and weight gradients.
//   trans_a = 0:  C[M,N] = act(A[M,K] * B[N,K]^T + bias)     (nn.Linear forward / input gradients)
//   trans_a = 1:  C[M,N] = A[K,M]^T * B[K,N] (+ C)           (weight gradients, contraction over rows)
// 64x64x16 tiles, 256 threads, 4x4 register tile per thread.
#include "decoder_common.cuh"

namespace zeggs {

template <bool TA>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb, const float* __restrict__ bias,
                                                    float* __restrict__ C, int ldc, int act, int accumulate) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    if (TA) {
      // A[k][m], B[k][n]: 16 x 64 each, 4 elements per thread, coalesced along m / n
      for (int i = tid; i < 16 * 64; i += 256) {
        int kk = i >> 6, mm = i & 63;
        int k = k0 + kk;
        As[kk][mm] = (k < K && m0 + mm < M) ? A[(size_t)k * lda + m0 + mm] : 0.f;
        Bs[kk][mm] = (k < K && n0 + mm < N) ? B[(size_t)k * ldb + n0 + mm] : 0.f;
      }
    } else {
      for (int i = tid; i < 16 * 64; i += 256) {
        int mm = i >> 4, kk = i & 15;
        int k = k0 + kk;
        As[kk][mm] = (k < K && m0 + mm < M) ? A[(size_t)(m0 + mm) * lda + k] : 0.f;
        Bs[kk][mm] = (k < K && n0 + mm < N) ? B[(size_t)(n0 + mm) * ldb + k] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      if (act == 1) v = elu_f(v); else if (act == 2) v = fmaxf(v, 0.f);
      if (accumulate) v += C[(size_t)m * ldc + n];
      C[(size_t)m * ldc + n] = v;
    }
  }
}

int sgemm_launch(int trans_a, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                 const float* bias, float* C, int ldc, int act, int accumulate, cudaStream_t stream) {
  ZCHECK_ARG(M > 0 && N > 0 && K > 0 && A && B && C, "sgemm: bad arguments M=%d N=%d K=%d", M, N, K);
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  if (trans_a) sgemm_kernel<true><<<grid, 256, 0, stream>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, act, accumulate);
  else sgemm_kernel<false><<<grid, 256, 0, stream>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, act, accumulate);
  count_launch();
  ZCHECK_LAUNCH();
  return ZEGGS_OK;
}

extern "C" int zeggs_sgemm(int trans_a, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                           const float* bias, float* C, int ldc, int act, int accumulate, void* stream) {
  return sgemm_launch(trans_a, M, N, K, A, lda, B, ldb, bias, C, ldc, act, accumulate, (cudaStream_t)stream);
}

}  // namespace zeggs
