// Register-resident small DFTs for the mel front end (csrc/mel.cu).  An 800-point complex FFT (= two real STFT frames,
// spectrograms.py:251-263) is split 32 x 25: every lane of a warp transforms 25 points in registers (dft25), the
// 32 x 25 intermediate is transposed through shared memory, and 25 lanes run a 32-point transform each (dft32).
// Both routines are __host__ __device__ so tests/host/mel_fft_host.cu can check them against numpy on the CPU.
#pragma once
#include <cuda_runtime.h>

namespace zeggs {

struct C2 { float re, im; };
__host__ __device__ inline C2 c2(float a, float b) { C2 r; r.re = a; r.im = b; return r; }
__host__ __device__ inline C2 operator+(C2 a, C2 b) { return c2(a.re + b.re, a.im + b.im); }
__host__ __device__ inline C2 operator-(C2 a, C2 b) { return c2(a.re - b.re, a.im - b.im); }
__host__ __device__ inline C2 cmul(C2 a, C2 b) { return c2(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__host__ __device__ inline C2 mul_mi(C2 a) { return c2(a.im, -a.re); }   // a * (-i)
__host__ __device__ inline C2 mul_pi(C2 a) { return c2(-a.im, a.re); }   // a * (+i)
__host__ __device__ inline C2 scale(float s, C2 a) { return c2(s * a.re, s * a.im); }

// forward 5-point DFT (sign -), in place
__host__ __device__ inline void dft5(C2& a0, C2& a1, C2& a2, C2& a3, C2& a4) {
  const float c1 = 0.30901699437494742f, c2_ = -0.80901699437494742f;
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const C2 t1 = a1 + a4, t2 = a2 + a3, t3 = a1 - a4, t4 = a2 - a3;
  const C2 m1 = a0 + scale(c1, t1) + scale(c2_, t2);
  const C2 m2 = a0 + scale(c2_, t1) + scale(c1, t2);
  const C2 n1 = scale(s1, t3) + scale(s2, t4);
  const C2 n2 = scale(s2, t3) - scale(s1, t4);
  a0 = a0 + t1 + t2;
  a1 = m1 + mul_mi(n1);
  a4 = m1 + mul_pi(n1);
  a2 = m2 + mul_mi(n2);
  a3 = m2 + mul_pi(n2);
}

// position p of dft25's output holds frequency k2 = p/5 + 5*(p%5)
__host__ __device__ constexpr int dft25_freq_of_pos(int p) { return p / 5 + 5 * (p % 5); }

// forward 25-point DFT of v[n2], n2 = 5a + b: five DFT5 over a, twiddle exp(-2 pi i b c / 25), five DFT5 over b.
// Result in place, permuted: v[5c + d] = X[c + 5d].
__host__ __device__ inline void dft25(C2 (&v)[25]) {
  // exp(-2 pi i m / 25), m = 1,2,3,4,6,8,9,12,16
  const C2 w1 = c2(0.96858316112863108f, -0.24868988716485479f), w2 = c2(0.87630668004386358f, -0.48175367410171532f);
  const C2 w3 = c2(0.72896862742141155f, -0.68454710592868873f), w4 = c2(0.53582679497899666f, -0.84432792550201508f);
  const C2 w6 = c2(0.06279051952931337f, -0.99802672842827156f), w8 = c2(-0.42577929156507272f, -0.90482705246601947f);
  const C2 w9 = c2(-0.63742398974868975f, -0.77051324277578925f), w12 = c2(-0.99211470131447788f, -0.12533323356430426f);
  const C2 w16 = c2(-0.63742398974868975f, 0.77051324277578925f);
#pragma unroll
  for (int b = 0; b < 5; ++b) dft5(v[b], v[5 + b], v[10 + b], v[15 + b], v[20 + b]);      // v[5c + b] = y_b[c]
  v[5 * 1 + 1] = cmul(v[5 * 1 + 1], w1);  v[5 * 1 + 2] = cmul(v[5 * 1 + 2], w2);  v[5 * 1 + 3] = cmul(v[5 * 1 + 3], w3);  v[5 * 1 + 4] = cmul(v[5 * 1 + 4], w4);
  v[5 * 2 + 1] = cmul(v[5 * 2 + 1], w2);  v[5 * 2 + 2] = cmul(v[5 * 2 + 2], w4);  v[5 * 2 + 3] = cmul(v[5 * 2 + 3], w6);  v[5 * 2 + 4] = cmul(v[5 * 2 + 4], w8);
  v[5 * 3 + 1] = cmul(v[5 * 3 + 1], w3);  v[5 * 3 + 2] = cmul(v[5 * 3 + 2], w6);  v[5 * 3 + 3] = cmul(v[5 * 3 + 3], w9);  v[5 * 3 + 4] = cmul(v[5 * 3 + 4], w12);
  v[5 * 4 + 1] = cmul(v[5 * 4 + 1], w4);  v[5 * 4 + 2] = cmul(v[5 * 4 + 2], w8);  v[5 * 4 + 3] = cmul(v[5 * 4 + 3], w12); v[5 * 4 + 4] = cmul(v[5 * 4 + 4], w16);
#pragma unroll
  for (int c = 0; c < 5; ++c) dft5(v[5 * c], v[5 * c + 1], v[5 * c + 2], v[5 * c + 3], v[5 * c + 4]);   // v[5c + d] = X[c + 5d]
}

// 5-bit reversal: dft32 leaves frequency k1 at position brev5(k1)
__host__ __device__ constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// forward 32-point DFT, radix-2 decimation in frequency, in place; v[brev5(k)] = X[k]
__host__ __device__ inline void dft32(C2 (&v)[32]) {
  // cos(2 pi j / 32), sin(2 pi j / 32), j = 0..7 (the rest by symmetry)
  const float cs[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                       0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f};
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int g = 0; g < (1 << s); ++g) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        const int i0 = g * 2 * half + j, i1 = i0 + half;
        const C2 a = v[i0], b = v[i1];
        v[i0] = a + b;
        const C2 d = a - b;
        const int m = j * (16 / half);          // twiddle exp(-2 pi i m / 32), m in [0, 16)
        if (m == 0) v[i1] = d;
        else if (m == 8) v[i1] = mul_mi(d);
        else if (m < 8) v[i1] = cmul(d, c2(cs[m], -cs[8 - m]));
        else v[i1] = cmul(d, c2(-cs[16 - m], -cs[m - 8]));
      }
    }
  }
}

}  // namespace zeggs
