"""Mel front end on the GPU -- host-side mirror of the reference's audio feature API.

Mirrors (same names / argument meaning):
  ZEGGS/audio/spectrograms.py:8-54   extract_mel_spectrogram_for_tts
  ZEGGS/data_pipeline.py:33-84       preprocess_audio
The arithmetic runs in `zeggs_mel_forward` (csrc/mel.cu); this file only builds the small
constant tables (Hann window, FFT twiddles, sparse Slaney filterbank) in float64 on the host
and moves buffers.  No CPU fallback.
"""
import math

import numpy as np
import torch

from . import _lib


def _hz_to_mel(f):  # spectrograms.py:446-473
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, log_step = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / log_step, f / f_sp)


def _mel_to_hz(m):  # spectrograms.py:476-503
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, log_step = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(log_step * (m - min_log_mel)), f_sp * m)


def mel_filterbank(n_fft, fs, n_mels, fmin, fmax, normalize=True):
    """Slaney filterbank [n_mels, n_fft//2+1] (spectrograms.py:386-443)."""
    if fmax is None:
        fmax = fs / 2.0
    nb = 1 + n_fft // 2
    fft_freqs = np.linspace(0, fs / 2.0, nb)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_freqs[None, :]
    w = np.zeros((n_mels, nb))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    if normalize:
        w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w


class MelFrontEnd:
    """Device-resident tables + launcher for one audio configuration (data_pipeline_conf.audio_conf)."""

    def __init__(self, device, sampling_rate=16000, filter_length=800, hop_length=200, n_mel_channels=80,
                 mel_fmin=20, mel_fmax=7600, min_clipping=1e-5, real_amplitude=True, normalize_mel_bins=True):
        self.device = torch.device(device)
        self.fs, self.n_fft, self.hop, self.n_mels = sampling_rate, filter_length, hop_length, n_mel_channels
        self.min_amp = float(min_clipping / (filter_length if real_amplitude else 1))   # spectrograms.py:86-88
        n = filter_length
        k = np.arange(n, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / (n - 1))                           # sps.hann(n), :230
        n2 = n // 2
        a1 = -2.0 * np.pi * np.arange(n2) / n2
        a2 = -2.0 * np.pi * np.arange(n2 + 1) / n
        tw = np.concatenate([np.stack([np.cos(a1), np.sin(a1)], 1), np.stack([np.cos(a2), np.sin(a2)], 1)], 0)
        fb = mel_filterbank(n, sampling_rate, n_mel_channels, mel_fmin, mel_fmax, normalize_mel_bins)
        start, length, off, ws = [], [], [], []
        for i in range(n_mel_channels):
            nz = np.nonzero(fb[i])[0]
            s, e = (int(nz[0]), int(nz[-1]) + 1) if len(nz) else (0, 0)
            start.append(s); length.append(e - s); off.append(len(ws)); ws.extend(fb[i, s:e].tolist())
        dev = self.device
        self.window = torch.tensor(window, dtype=torch.float32, device=dev)
        self.twiddle = torch.tensor(tw, dtype=torch.float32, device=dev).contiguous()
        self.fb_start = torch.tensor(start, dtype=torch.int32, device=dev)
        self.fb_len = torch.tensor(length, dtype=torch.int32, device=dev)
        self.fb_off = torch.tensor(off, dtype=torch.int32, device=dev)
        self.fb_w = torch.tensor(ws if ws else [0.0], dtype=torch.float32, device=dev)

    def num_frames(self, n_samples):
        return _lib.lib().zeggs_mel_num_frames(int(n_samples), self.n_fft, self.hop)

    def forward(self, wav, anim_fs=None, anim_length=None, want_mel=False, want_feat=True):
        """wav [N, n_samples] f32 CUDA -> (mel [N, n_mels, L] or None, feat [N, anim_length, n_mels+1] or None)."""
        if wav.dim() == 1:
            wav = wav[None]
        wav = wav.contiguous().float()
        N, ns = wav.shape
        L = self.num_frames(ns)
        mel = torch.empty((N, self.n_mels, L), dtype=torch.float32, device=wav.device) if want_mel else None
        feat = None
        fpa = 0.0
        if want_feat:
            fpa = (self.fs / self.hop) / anim_fs                                           # data_pipeline.py:68
            feat = torch.empty((N, anim_length, self.n_mels + 1), dtype=torch.float32, device=wav.device)
        a = _lib.MelArgs(n_clips=N, n_samples=ns, n_fft=self.n_fft, hop=self.hop, n_mels=self.n_mels,
                         anim_length=int(anim_length or 0), min_amp=self.min_amp, frames_per_anim=float(fpa),
                         wav=_lib.ptr(wav), window=_lib.ptr(self.window), twiddle=_lib.ptr(self.twiddle),
                         fb_start=_lib.ptr(self.fb_start), fb_len=_lib.ptr(self.fb_len), fb_off=_lib.ptr(self.fb_off),
                         fb_w=_lib.ptr(self.fb_w), mel_out=_lib.ptr(mel), feat_out=_lib.ptr(feat), fb_total=int(self.fb_w.numel()))
        _lib.check(_lib.lib().zeggs_mel_forward(a, _lib.stream_ptr()), "zeggs_mel_forward")
        return mel, feat


_cache = {}


def _front_end(device, **kw):
    key = (str(device),) + tuple(sorted(kw.items()))
    if key not in _cache:
        _cache[key] = MelFrontEnd(device, **kw)
    return _cache[key]


def _conf_kwargs(params):
    g = (lambda k: params[k]) if isinstance(params, dict) else (lambda k: getattr(params, k))
    if g("pre_emphasis"):
        raise _lib.ZeggsError("pre_emphasis=True is not on the accelerated path (shipped confs use False)")
    if not (g("centered") and g("normalize_range") and g("resample_method") == "linear"):
        raise _lib.ZeggsError("only centered / normalize_range / linear resampling (the shipped audio_conf) is supported")
    return dict(sampling_rate=g("sampling_rate"), filter_length=g("filter_length"), hop_length=g("hop_length"),
                n_mel_channels=g("n_mel_channels"), mel_fmin=g("mel_fmin"), mel_fmax=g("mel_fmax"),
                min_clipping=g("min_clipping"), real_amplitude=g("real_amplitude"),
                normalize_mel_bins=g("normalize_mel_bins"))


def preprocess_audio(audio_data, anim_fs, anim_length, params, feature_type, device="cuda"):
    """Drop-in for data_pipeline.preprocess_audio (data_pipeline.py:33-84): numpy/torch [T] (or [N,T]) in,
    float32 [anim_length, 81] (numpy for numpy input, CUDA tensor for tensor input) out.
    Loudness normalisation (pyloudnorm, :34-39) is a scalar gain outside this path: apply it before."""
    nl = params["normalize_loudness"] if isinstance(params, dict) else getattr(params, "normalize_loudness", False)
    if nl:
        raise _lib.ZeggsError("normalize_loudness=True: apply the BS.1770 gain before calling (not on this path)")
    if list(feature_type) != ["mel_spec", "energy"]:
        raise _lib.ZeggsError("feature_type must be ['mel_spec', 'energy'] (the shipped audio_feature_type)")
    as_numpy = isinstance(audio_data, np.ndarray)
    wav = torch.as_tensor(audio_data, dtype=torch.float32)
    if not wav.is_cuda:
        wav = wav.pin_memory().to(device, non_blocking=True) if as_numpy else wav.to(device)
    fe = _front_end(wav.device, **_conf_kwargs(params))
    _, feat = fe.forward(wav, anim_fs, anim_length, want_mel=False, want_feat=True)
    if wav.dim() == 1 or (as_numpy and np.ndim(audio_data) == 1):
        feat = feat[0]
    return feat.cpu().numpy() if as_numpy else feat


def extract_mel_spectrogram_for_tts(wav_signal, fs, n_fft, step_size, n_mels, mel_fmin, mel_fmax, min_amplitude,
                                    pre_emphasis=True, pre_emph_coeff=0.97, dynamic_range=None, real_amplitude=True,
                                    centered=True, normalize_mel_bins=True, normalize_range=True, logger=None,
                                    device="cuda"):
    """Drop-in for spectrograms.extract_mel_spectrogram_for_tts (spectrograms.py:8-54) -> (mel[n_mels, L], wav)."""
    if pre_emphasis or dynamic_range or not (centered and normalize_range and min_amplitude):
        raise _lib.ZeggsError("unsupported option combination (accelerated path = the shipped audio_conf)")
    as_numpy = isinstance(wav_signal, np.ndarray)
    wav = torch.as_tensor(wav_signal, dtype=torch.float32).to(device)
    fe = _front_end(wav.device, sampling_rate=fs, filter_length=n_fft, hop_length=step_size, n_mel_channels=n_mels,
                    mel_fmin=mel_fmin, mel_fmax=mel_fmax, min_clipping=min_amplitude, real_amplitude=real_amplitude,
                    normalize_mel_bins=normalize_mel_bins)
    mel, _ = fe.forward(wav, want_mel=True, want_feat=False)
    mel = mel[0]
    return (mel.cpu().numpy() if as_numpy else mel), wav_signal
