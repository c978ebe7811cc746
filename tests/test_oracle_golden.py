"""The oracle restatement (oracle/) against the committed golden vectors, which were produced
by the unmodified reference (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle, model_oracle as mo
from zeggs_b200 import synth

NAMES = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt"]


def tt(d):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in d.items()}


@pytest.mark.parametrize("hop", [200, 160])
def test_mel_oracle_matches_reference_golden(golden_dir, hop):
    g = np.load(os.path.join(golden_dir, "mel_small.npz"))
    wav = g["wav"]
    for i in range(wav.shape[0]):
        mel = mel_oracle.mel_spectrogram(wav[i], hop=hop)
        assert mel.shape == g[f"mel_hop{hop}"][i].shape
        # float64 restatement: batched layout only, same arithmetic
        assert np.max(np.abs(mel - g[f"mel_hop{hop}"][i])) <= 1e-12
        n60 = g[f"feat_hop{hop}"].shape[1]
        feat = mel_oracle.preprocess_audio(wav[i], 60, n60, hop=hop)
        assert feat.dtype == np.float32 and feat.shape == (n60, 81)
        assert np.max(np.abs(feat - g[f"feat_hop{hop}"][i])) <= 1e-6


def test_num_frames_rule():
    # spectrograms.py:242-245: 10 s @ hop 200 -> 800 frames, hop 160 -> 1000
    assert mel_oracle.num_frames(160000, 800, 200) == 800
    assert mel_oracle.num_frames(160000, 800, 160) == 1000
    assert mel_oracle.num_frames(16001, 800, 200) == 1 + (16001 + 800 - 800) // 200


def _run_oracle(g):
    H, B, T, T_ex = int(g["H"]), int(g["B"]), int(g["T"]), int(g["T_ex"])
    P = tt(synth.make_params(H=H, seed=int(g["param_seed"])))
    for v in P.values():
        v.requires_grad_(True)
    st = synth.load_stats()
    f = lambda k: torch.as_tensor(st[k], dtype=torch.float32)
    win = tt(synth.make_pose_windows(B, T, seed=int(g["input_seed"])))
    audio = torch.from_numpy(synth.make_audio_features(B, T, seed=int(g["input_seed"])))
    style_ex = torch.from_numpy(synth.make_style_example(B, T_ex, seed=int(g["input_seed"])))
    speech = mo.speech_encoder(P, (audio - f("audio_input_mean")) / f("audio_input_std"))
    z, mu, logvar = mo.style_encoder(P, (style_ex - f("anim_input_mean")) / f("anim_input_std"),
                                     eps=torch.from_numpy(g["eps"]))
    O = mo.decoder_forward(P, *[win[n][:, 0] for n in NAMES], win["gaze_pos"], speech,
                           z.unsqueeze(1).repeat(1, T, 1), f("anim_input_mean"), f("anim_input_std"),
                           f("anim_output_mean"), f("anim_output_std"), float(st["dt"]))
    loss, terms = mo.train_losses(O, [win[n] for n in NAMES], win["gaze_pos"], st["parents"], float(st["dt"]),
                                  mu, logvar, int(g["iteration"]))
    return P, speech, (z, mu, logvar), O, loss, terms


@pytest.mark.parametrize("tag", ["h64", "h128", "h384", "h1024"])
def test_network_oracle_matches_reference_golden(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"train_{tag}.npz"))
    P, speech, (z, mu, logvar), O, loss, terms = _run_oracle(g)
    assert np.max(np.abs(speech.detach().numpy() - g["speech"])) <= 2e-6
    assert np.max(np.abs(mu.detach().numpy() - g["mu"])) <= 5e-6
    assert np.max(np.abs(logvar.detach().numpy() - g["logvar"])) <= 5e-6
    assert np.max(np.abs(z.detach().numpy() - g["z"])) <= 5e-6
    for n, o in zip(NAMES, O):
        ref = g["O_" + n]
        tol = 2e-5 * max(1.0, float(np.max(np.abs(ref))))
        assert np.max(np.abs(o.detach().numpy() - ref)) <= tol, n
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    for k, v in terms.items():
        assert abs(float(v) - float(g["loss_" + k])) <= 2e-5 * max(1e-3, abs(float(g["loss_" + k]))), k
    # backward: gradient of every parameter on the path
    names = [k for k in P.keys()]
    grads = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    for k, gr in zip(names, grads):
        ref_n = float(g["gradnorm." + k])
        assert gr is not None, k
        assert abs(float(gr.double().norm()) - ref_n) <= 2e-4 * max(ref_n, 1e-6), k
        if "grad." + k in g.files:
            assert np.max(np.abs(gr.numpy() - g["grad." + k])) <= 2e-4 * max(1e-6, float(np.max(np.abs(g["grad." + k])))), k


def test_radam_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "radam.npz"))
    p = torch.from_numpy(g["p0"].copy())
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for i in range(g["grads"].shape[0]):
        mo.radam_step(p, torch.from_numpy(g["grads"][i]), m, v, i + 1, lr=1e-4, eps=1e-5)
        assert np.max(np.abs(p.numpy() - g["traj"][i])) <= 1e-7, i


def test_oracle_on_shipped_v1_weights_matches_reference_golden(golden_dir):
    """The v1 pickles' weights (tests/_v1/weights.npz, git-ignored, written by oracle/make_golden.py) through the oracle against
    the outputs the unmodified reference produced from the pickles themselves (tests/golden/v1_pretrained.npz)."""
    wpath = os.path.join(os.path.dirname(golden_dir), "_v1", "weights.npz")
    if not os.path.exists(wpath):
        pytest.skip("tests/_v1/weights.npz not present (python -m oracle.make_golden v1)")
    g = np.load(os.path.join(golden_dir, "v1_pretrained.npz"))
    P = tt(dict(np.load(wpath)))
    B, T, T_ex, seed = int(g["B"]), int(g["T"]), int(g["T_ex"]), int(g["input_seed"])
    st = synth.load_stats()
    f = lambda k: torch.as_tensor(st[k], dtype=torch.float32)
    win = tt(synth.make_pose_windows(B, T, seed=seed))
    audio = torch.from_numpy(synth.make_audio_features(B, T, seed=seed))
    style_ex = torch.from_numpy(synth.make_style_example(B, T_ex, seed=seed))
    with torch.no_grad():
        speech = mo.speech_encoder(P, (audio - f("audio_input_mean")) / f("audio_input_std"))
        z, mu, logvar = mo.style_encoder(P, (style_ex - f("anim_input_mean")) / f("anim_input_std"), eps=torch.zeros(B, 64))
        O = mo.decoder_forward(P, *[win[n][:, 0] for n in NAMES], win["gaze_pos"], speech, mu.unsqueeze(1).repeat(1, T, 1),
                               f("anim_input_mean"), f("anim_input_std"), f("anim_output_mean"), f("anim_output_std"), float(st["dt"]))
    assert np.max(np.abs(speech.numpy() - g["speech"])) <= 1e-5
    assert np.max(np.abs(mu.numpy() - g["mu"])) <= 2e-5
    for n, o in zip(NAMES, O):
        ref = g["O_" + n]
        assert np.max(np.abs(o.numpy() - ref)) <= 1e-4 * max(1.0, float(np.max(np.abs(ref)))), n
