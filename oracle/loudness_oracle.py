"""CPU restatement of the loudness normalisation the reference applies ahead of the mel front end
(ZEGGS/data_pipeline.py:34-39: `pyln.Meter(rate).integrated_loudness(x)` then `pyln.normalize.loudness(x, L, -20.0)`).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu legs).

**Parity unpinned.**  The arithmetic lives in the third-party package `pyloudnorm==0.1.0` (requirements.txt), which is NOT
installed in this image and is not vendored under /root/reference, and the reference holds no test or golden vector for it.
This file restates the published algorithm of that version (ITU-R BS.1770-4 as implemented by pyloudnorm 0.1.0:
`Meter.__init__`, `IIRfilter.generate_coefficients`, `Meter.integrated_loudness`, `normalize.loudness`):

  * K-weighting = two biquads applied in this order: high shelf (G = +4 dB, Q = 1/sqrt(2), fc = 1500 Hz) and high pass
    (G = 0 dB, Q = 0.5, fc = 38 Hz), RBJ-cookbook coefficient formulas, `scipy.signal.lfilter` (float64 recursion).  The
    filtered signal is written back into a copy of the INPUT array, i.e. rounded to the input dtype (float32 for the
    reference's `read_wavfile(..., out_type='float32')`) after each stage.
  * gating blocks of T_g = 0.4 s, 75 % overlap: block j covers samples [int(T_g*(j*0.25)*rate), int(T_g*(j*0.25+1)*rate)),
    numBlocks = int(round((T - T_g) / (T_g*0.25)) + 1); z_j = sum(x^2) / (T_g*rate); l_j = -0.691 + 10 log10(z_j) (mono, G = 1)
  * absolute gate -70 LUFS, relative gate = loudness of the absolutely-gated mean - 10 LU, integrated loudness =
    -0.691 + 10 log10(mean z over blocks above both gates)
  * gain = 10^((target - loudness) / 20), output = gain * x.
"""
import numpy as np
from scipy import signal

T_G, OVERLAP, GAMMA_A = 0.4, 0.75, -70.0


def biquad(G, Q, fc, rate, kind):
    """pyloudnorm 0.1.0 IIRfilter.generate_coefficients -> (b[3], a[3]) normalised by a0."""
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    if kind == "high_shelf":
        b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
        b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
        b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
        a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
        a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
        a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
    elif kind == "high_pass":
        b0 = (1 + np.cos(w0)) / 2
        b1 = -(1 + np.cos(w0))
        b2 = (1 + np.cos(w0)) / 2
        a0 = 1 + alpha
        a1 = -2 * np.cos(w0)
        a2 = 1 - alpha
    else:
        raise ValueError(kind)
    return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0


def k_weighting(rate):
    return [biquad(4.0, 1.0 / np.sqrt(2.0), 1500.0, rate, "high_shelf"), biquad(0.0, 0.5, 38.0, rate, "high_pass")]


def block_bounds(n_samples, rate):
    """(l_j, u_j) of every gating block, computed with the same float64 expressions as the package."""
    step = 1.0 - OVERLAP
    T = n_samples / rate
    nb = int(np.round(((T - T_G) / (T_G * step))) + 1)
    lo = [int(T_G * (j * step) * rate) for j in range(nb)]
    hi = [int(T_G * (j * step + 1) * rate) for j in range(nb)]
    return lo, hi


def integrated_loudness(x, rate):
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError("mono audio expected (the reference passes a 1-D array)")
    if x.shape[0] < T_G * rate:
        raise ValueError("Audio must have length greater than the block size.")
    y = x.copy()
    for b, a in k_weighting(rate):
        y[:] = signal.lfilter(b, a, y)          # rounds to the input dtype, like the package's in-place channel assignment
    lo, hi = block_bounds(len(x), rate)
    z = np.array([(1.0 / (T_G * rate)) * np.sum(np.square(y[l:u])) for l, u in zip(lo, hi)], dtype=np.float64)
    with np.errstate(divide="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
    J = [j for j in range(len(z)) if l[j] >= GAMMA_A]
    with np.errstate(divide="ignore", invalid="ignore"):
        z_avg = np.mean(z[J]) if J else np.nan
        gamma_r = -0.691 + 10.0 * np.log10(z_avg) - 10.0
        J = [j for j in range(len(z)) if (l[j] > gamma_r and l[j] > GAMMA_A)]
        z_avg = np.nan_to_num(np.mean(z[J]) if J else np.nan)
        return float(-0.691 + 10.0 * np.log10(z_avg))


def loudness_gain(x, rate, target=-20.0):
    return float(np.power(10.0, (target - integrated_loudness(x, rate)) / 20.0))


def normalize_loudness(x, rate, target=-20.0):
    """What data_pipeline.py:34-39 hands to the mel front end (float64 array: python-float gain x float32 samples)."""
    return loudness_gain(x, rate, target) * np.asarray(x)
