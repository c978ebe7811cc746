// Host-side harness: runs the SAME __host__ __device__ small-DFT routines (csrc/mel_fft.cuh) the mel kernel keeps in
// registers, emulating the warp-level 32 x 25 decomposition lane by lane (tests/test_mel_fft_host.py).  Test tooling only.
#include <cmath>
#include "../../ubisoft-laforge-zeroeggs_b200/csrc/mel_fft.cuh"
using namespace zeggs;

extern "C" void dft25_host(const float* in, float* out) {      // in/out: 25 complex (re, im interleaved), natural order
  C2 v[25];
  for (int i = 0; i < 25; ++i) v[i] = c2(in[2 * i], in[2 * i + 1]);
  dft25(v);
  for (int p = 0; p < 25; ++p) { const int k = dft25_freq_of_pos(p); out[2 * k] = v[p].re; out[2 * k + 1] = v[p].im; }
}
extern "C" void dft32_host(const float* in, float* out) {
  C2 v[32];
  for (int i = 0; i < 32; ++i) v[i] = c2(in[2 * i], in[2 * i + 1]);
  dft32(v);
  for (int k = 0; k < 32; ++k) { out[2 * k] = v[brev5(k)].re; out[2 * k + 1] = v[brev5(k)].im; }
}
// two real 800-sample frames -> magnitudes of rfft bins 0..400 of each, exactly as the kernel's lanes do it
extern "C" void two_frames_host(const float* fa, const float* fb, float* amp_a, float* amp_b) {
  static C2 buf[25][32];
  for (int lane = 0; lane < 32; ++lane) {
    C2 v[25];
    for (int j = 0; j < 25; ++j) v[j] = c2(fa[lane + 32 * j], fb[lane + 32 * j]);
    dft25(v);
    for (int p = 0; p < 25; ++p) {
      const int k2 = dft25_freq_of_pos(p);
      const double ang = -2.0 * M_PI * (double)((lane * k2) % 800) / 800.0;
      buf[k2][lane] = k2 ? cmul(v[p], c2((float)cos(ang), (float)sin(ang))) : v[p];
    }
  }
  static C2 u[25][32];
  for (int lane = 0; lane < 25; ++lane) {
    for (int n1 = 0; n1 < 32; ++n1) u[lane][n1] = buf[lane][n1];
    dft32(u[lane]);
  }
  for (int lane = 0; lane < 25; ++lane)
    for (int k1 = 0; k1 <= 16; ++k1) {
      const int k = lane + 25 * k1;
      if (k > 400) continue;
      const C2 P = u[lane][brev5(k1)];
      const C2 Q = lane == 0 ? u[0][brev5((32 - k1) & 31)] : u[25 - lane][brev5(31 - k1)];
      const float are = 0.5f * (P.re + Q.re), aim = 0.5f * (P.im - Q.im), bre = 0.5f * (P.im + Q.im), bim = -0.5f * (P.re - Q.re);
      amp_a[k] = sqrtf(are * are + aim * aim);
      amp_b[k] = sqrtf(bre * bre + bim * bim);
    }
}
