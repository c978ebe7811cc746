// Mel front end: reflect-pad -> symmetric-Hann STFT magnitude / n_fft -> Slaney mel filterbank -> clip ->
// dB -> [0,1] -> (optional) ln(10^(s/20)), 80->60 fps linear resample, energy channel.
// Replaces ZEGGS/audio/spectrograms.py:216-269, :161-183, :386-503, :57-131 and
// ZEGGS/data_pipeline.py:62-82 (one fused kernel; HBM-bound: each sample is read ~once, each output
// written once).
//
// One CTA = one clip x a tile of 32 consecutive STFT frames (+1 neighbour frame for the resampler).
// The tile's samples are staged once in shared memory (coalesced, reflect padding resolved at load);
// each warp then transforms one frame at a time: the 800 windowed real samples are packed into a
// 400-point complex sequence, transformed by a radix 4,4,5,5 Stockham FFT in shared memory, split
// into the 401 real-FFT bins, and reduced through the sparse mel filterbank.
#include "decoder_common.cuh"

namespace zeggs {

constexpr int MEL_FT = 32;       // frames per tile
constexpr int MEL_N2 = 400;      // complex FFT length (n_fft / 2)
constexpr int MEL_WARPS = 8;

struct C2 { float re, im; };
__device__ __forceinline__ C2 c2(float a, float b) { C2 r; r.re = a; r.im = b; return r; }
__device__ __forceinline__ C2 operator+(C2 a, C2 b) { return c2(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ C2 operator-(C2 a, C2 b) { return c2(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ C2 cmul(C2 a, C2 b) { return c2(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ C2 mul_mi(C2 a) { return c2(a.im, -a.re); }   // a * (-i)
__device__ __forceinline__ C2 mul_pi(C2 a) { return c2(-a.im, a.re); }   // a * (+i)
__device__ __forceinline__ C2 scale(float s, C2 a) { return c2(s * a.re, s * a.im); }

// one Stockham DIF stage of radix R on a length-400 sequence: n = current sub-length, s = stride (n*s = 400)
template <int R>
__device__ __forceinline__ void fft_stage(const C2* __restrict__ x, C2* __restrict__ y, int n, int s,
                                          const float2* __restrict__ tw, int lane) {
  const int m = n / R;
  for (int bfly = lane; bfly < MEL_N2 / R; bfly += 32) {
    const int p = bfly / s, q = bfly - p * s;
    C2 a[R];
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = x[q + s * (p + j * m)];
    C2 b[R];
    if (R == 4) {
      C2 t0 = a[0] + a[2], t1 = a[0] - a[2], t2 = a[1] + a[3], t3 = a[1] - a[3];
      b[0] = t0 + t2;
      b[1] = t1 + mul_mi(t3);
      b[2] = t0 - t2;
      b[3] = t1 + mul_pi(t3);
    } else {
      const float c1 = 0.30901699437494742f, c2_ = -0.80901699437494742f;
      const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
      C2 t1 = a[1] + a[4], t2 = a[2] + a[3], t3 = a[1] - a[4], t4 = a[2] - a[3];
      b[0] = a[0] + t1 + t2;
      C2 m1 = a[0] + scale(c1, t1) + scale(c2_, t2);
      C2 m2 = a[0] + scale(c2_, t1) + scale(c1, t2);
      C2 n1 = scale(s1, t3) + scale(s2, t4);
      C2 n2 = scale(s2, t3) - scale(s1, t4);
      b[1] = m1 + mul_mi(n1);
      b[4] = m1 + mul_pi(n1);
      b[2] = m2 + mul_mi(n2);
      b[3] = m2 + mul_pi(n2);
    }
    y[q + s * (R * p)] = b[0];
#pragma unroll
    for (int k = 1; k < R; ++k) {
      const float2 w = tw[p * k * s];   // exp(-2 pi i p k / n) = exp(-2 pi i p k s / 400)
      y[q + s * (R * p + k)] = cmul(b[k], c2(w.x, w.y));
    }
  }
}

__global__ void __launch_bounds__(MEL_WARPS * 32) mel_kernel(zeggs_mel_args a, int L, double fpa) {
  extern __shared__ __align__(16) float sm[];
  const int n_fft = a.n_fft, hop = a.hop, nm = a.n_mels;
  const int nsamp_tile = MEL_FT * hop + n_fft;                 // frames f0..f0+32 inclusive need (32*hop + n_fft)
  float* samp = sm;                                            // [nsamp_tile]
  float* melt = samp + ((nsamp_tile + 3) & ~3);                // [(FT+1)][nm]  log-mel m = s*ln10/20
  float* ener = melt + (MEL_FT + 1) * nm;                      // [(FT+1)]
  float* sdb = ener + ((MEL_FT + 1 + 3) & ~3);                 // [(FT+1)][nm]  s in [0,1] (only if mel_out)
  C2* fbuf = reinterpret_cast<C2*>(sdb + (a.mel_out ? (MEL_FT + 1) * nm : 0));   // [warps][2][400]
  const int clip = blockIdx.y, f0 = blockIdx.x * MEL_FT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* wav = a.wav + (size_t)clip * a.n_samples;
  const int Ts = a.n_samples, Te = Ts < n_fft ? n_fft : Ts;    // spectrograms.py:233-234 zero-extends short clips
  // ---- stage the tile's samples (reflect padding, spectrograms.py:237-239)
  for (int i = tid; i < nsamp_tile; i += blockDim.x) {
    int o = f0 * hop + i - n_fft / 2;
    if (o < 0) o = -o;
    if (o >= Te) o = 2 * (Te - 1) - o;
    float v = 0.f;
    if (o >= 0 && o < Ts) v = __ldg(wav + o);
    samp[i] = v;
  }
  __syncthreads();
  const float2* tw = reinterpret_cast<const float2*>(a.twiddle);            // [400]
  const float2* tw2 = tw + MEL_N2;                                          // [401]  exp(-2 pi i k / 800)
  C2* A = fbuf + warp * 2 * MEL_N2;
  C2* Bf = A + MEL_N2;
  const float inv_nfft = 1.0f / (float)n_fft;
  const float R_db = -20.0f * log10f(a.min_amp);                            // spectrograms.py:127
  for (int fl = warp; fl <= MEL_FT; fl += MEL_WARPS) {
    const int f = f0 + fl;
    if (f >= L) break;
    const float* xs = samp + fl * hop;
    for (int m = lane; m < MEL_N2; m += 32)
      A[m] = c2(xs[2 * m] * __ldg(a.window + 2 * m), xs[2 * m + 1] * __ldg(a.window + 2 * m + 1));
    __syncwarp();
    fft_stage<4>(A, Bf, 400, 1, tw, lane); __syncwarp();
    fft_stage<4>(Bf, A, 100, 4, tw, lane); __syncwarp();
    fft_stage<5>(A, Bf, 25, 16, tw, lane); __syncwarp();
    fft_stage<5>(Bf, A, 5, 80, tw, lane); __syncwarp();
    // real-FFT split + magnitude / n_fft -> amp[0..400] (reuse Bf as float storage)
    float* amp = reinterpret_cast<float*>(Bf);
    for (int k = lane; k <= MEL_N2; k += 32) {
      C2 zk = A[k == MEL_N2 ? 0 : k];
      C2 zc = A[k == 0 ? 0 : MEL_N2 - k];
      zc.im = -zc.im;
      C2 e = scale(0.5f, zk + zc);
      C2 o = mul_mi(scale(0.5f, zk - zc));
      const float2 w = tw2[k];
      C2 xk = e + cmul(o, c2(w.x, w.y));
      amp[k] = sqrtf(xk.re * xk.re + xk.im * xk.im) * inv_nfft;
    }
    __syncwarp();
    float esum = 0.f;
    for (int band = lane; band < nm; band += 32) {
      const int st = a.fb_start[band], ln = a.fb_len[band];
      const float* wgt = a.fb_w + a.fb_off[band];
      float acc = 0.f;
      for (int i = 0; i < ln; ++i) acc = fmaf(__ldg(wgt + i), amp[st + i], acc);
      float v = fmaxf(fabsf(acc), a.min_amp);                               // spectrograms.py:110-116
      float s = (20.0f * log10f(v) + R_db) / R_db;                          // :119, :127-129
      float mm = s * 0.11512925464970229f;                                  // ln(10^(s/20)), data_pipeline.py:62-63
      melt[fl * nm + band] = mm;
      if (a.mel_out) sdb[fl * nm + band] = s;
      float ex = expf(mm);
      esum = fmaf(ex, ex, esum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, o);
    if (lane == 0) ener[fl] = sqrtf(esum);                                  // data_pipeline.py:28-30
    __syncwarp();
  }
  __syncthreads();
  // ---- mel_out[clip][band][f] (frames contiguous)
  if (a.mel_out) {
    for (int i = tid; i < nm * MEL_FT; i += blockDim.x) {
      int band = i / MEL_FT, fl = i % MEL_FT;
      if (f0 + fl < L) a.mel_out[((size_t)clip * nm + band) * L + f0 + fl] = sdb[fl * nm + band];
    }
  }
  // ---- 60 fps resample (data_pipeline.py:65-82): rows k with clamp(floor(k*fpa), 0, L-2) in [f0, f0+32)
  if (a.feat_out) {
    int k_lo = (int)floor((double)f0 / fpa) - 1;
    int k_hi = (int)ceil((double)(f0 + MEL_FT) / fpa) + 1;
    if (k_lo < 0) k_lo = 0;
    if (k_hi > a.anim_length) k_hi = a.anim_length;
    const int nch = nm + 1;
    for (int i = tid; i < (k_hi - k_lo) * nch; i += blockDim.x) {
      const int k = k_lo + i / nch, ch = i % nch;
      const double pos = (double)k * fpa;
      int i0 = (int)floor(pos);
      if (i0 > L - 2) i0 = L - 2;
      if (i0 < 0) i0 = 0;
      if (i0 < f0 || i0 >= f0 + MEL_FT) continue;
      const float wgt = (float)(pos - (double)i0);
      const int fl = i0 - f0;
      float v0, v1;
      if (ch < nm) { v0 = melt[fl * nm + ch]; v1 = (L > 1) ? melt[(fl + 1) * nm + ch] : v0; }
      else { v0 = ener[fl]; v1 = (L > 1) ? ener[fl + 1] : v0; }
      a.feat_out[((size_t)clip * a.anim_length + k) * nch + ch] = v0 + (v1 - v0) * wgt;
    }
  }
}

extern "C" int zeggs_mel_num_frames(int n_samples, int n_fft, int hop) {
  // spectrograms.py:233-245 (centered)
  if (n_samples < 1 || n_fft < 2 || hop < 1) return -1;
  long long n = n_samples < n_fft ? n_fft : n_samples;
  n += 2 * (n_fft / 2);
  long long L = (n - n_fft) / hop;
  if (n % hop != 0) L += 1;
  return (int)L;
}

extern "C" int zeggs_mel_forward(const zeggs_mel_args* ap, void* stream_) {
  ZCHECK_ARG(ap != nullptr, "mel: null args");
  const zeggs_mel_args& a = *ap;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (a.n_fft != 2 * MEL_N2) { set_error("mel: n_fft=%d unsupported (this build transforms n_fft=800 only)", a.n_fft); return ZEGGS_ERR_UNSUPPORTED; }
  ZCHECK_ARG(a.n_clips >= 0 && a.n_samples >= 1 && a.hop >= 1 && a.hop <= 800 && a.n_mels >= 1 && a.n_mels <= 256, "mel: bad shape");
  ZCHECK_ARG(a.wav && a.window && a.twiddle && a.fb_start && a.fb_len && a.fb_off && a.fb_w, "mel: null table/input pointer");
  ZCHECK_ARG(a.mel_out || a.feat_out, "mel: no output requested");
  ZCHECK_ARG(a.n_samples >= a.n_fft / 2 + 1, "mel: clip shorter than n_fft/2+1 samples cannot be reflect-padded");
  if (a.n_clips == 0) return ZEGGS_OK;
  const int L = zeggs_mel_num_frames(a.n_samples, a.n_fft, a.hop);
  ZCHECK_ARG(L >= 1, "mel: no frames");
  if (a.feat_out) {
    ZCHECK_ARG(a.anim_length >= 1 && a.frames_per_anim > 0.0, "mel: bad resample parameters");
    ZCHECK_ARG((double)(a.anim_length - 1) * a.frames_per_anim <= (double)(L - 1) + 1e-9,
               "mel: anim_length %d needs frames beyond L=%d (reference would produce NaN, data_pipeline.py:473)", a.anim_length, L);
  }
  const int nsamp_tile = MEL_FT * a.hop + a.n_fft;
  size_t smem = (size_t)(((nsamp_tile + 3) & ~3) + (MEL_FT + 1) * a.n_mels * (a.mel_out ? 2 : 1) + ((MEL_FT + 1 + 3) & ~3)) * sizeof(float)
              + (size_t)MEL_WARPS * 2 * MEL_N2 * sizeof(C2);
  ZCHECK_CUDA(cudaFuncSetAttribute(mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(L, MEL_FT), a.n_clips);
  ScopedTimer tm("mel", stream);
  mel_kernel<<<grid, MEL_WARPS * 32, smem, stream>>>(a, L, a.frames_per_anim);
  count_launch();
  ZCHECK_LAUNCH();
  return ZEGGS_OK;
}

}  // namespace zeggs
