"""CPU check of the algebra behind the tensor-core engine's folded recurrence (csrc/decoder_fwd_tc.cu, decoder_bwd_tc.cu):
the input GEMM of step t+1 applied to vectorize_input(devectorize_output(layer2(h1(t)))) (modules.py:713, :728, :172-175) equals
Mfold h1(t) + cfold + Wx[:, 1131:1134] gaze(t+1) with Mfold = (Wx[:, :1131] diag(os/is)) W2 -- in float64, via the oracle's own
vectorize / devectorize functions.  No GPU involved."""
import numpy as np
import torch

from oracle import model_oracle as mo
from tests._util import stats_tensors
from zeggs_b200 import synth


def test_layer2_folds_into_the_next_input_gemm():
    H, B = 64, 5
    st = stats_tensors()
    f64 = lambda t: t.double()
    im, is_, om, os_ = (f64(st[k]) for k in ("anim_input_mean", "anim_input_std", "anim_output_mean", "anim_output_std"))
    P = {k: torch.from_numpy(v).double() for k, v in synth.make_params(H=H, seed=5, with_style=False).items()}
    pre = "decoder.recurrent_decoder."
    W0, Wih0, W2, b2 = P[pre + "layer0.weight"], P[pre + "layer1.weight_ih_l0"], P[pre + "layer2.weight"], P[pre + "layer2.bias"]
    Wx = torch.cat([W0[:, :1134], Wih0[:, H:H + 1134]], 0)                      # [4H, 1134]: pose columns of layer0 / GRU0
    rs = np.random.RandomState(1)
    h1 = torch.from_numpy(rs.randn(B, H))
    win = {k: torch.from_numpy(v).double() for k, v in synth.make_pose_windows(B, 2, seed=9).items()}
    root_pos, root_rot, gaze_next = win["root_pos"][:, 0], win["root_rot"][:, 0], win["gaze_pos"][:, 1]
    dt = float(st["dt"])
    # the reference's two-step route: layer2 -> devectorize (de-normalise, integrate the root) -> vectorize (gaze, normalise)
    y = h1 @ W2.T + b2
    out = mo.devectorize_output(y, root_pos, root_rot, dt, om, os_)
    x_next = mo.vectorize_input(*out, gaze_next, im, is_)                       # [B, 1134]
    s1_ref = x_next @ Wx.T
    # the folded route
    D = (os_ / is_[:1131])
    Mfold = (Wx[:, :1131] * D) @ W2                                             # [4H, H]
    cfold = Wx[:, :1131] @ ((b2 * os_ + om - im[:1131]) / is_[:1131])
    gz = x_next[:, 1131:1134]                                                   # normalised gaze direction of step t+1
    s1_fold = h1 @ Mfold.T + cfold + gz @ Wx[:, 1131:1134].T
    assert float((s1_fold - s1_ref).abs().max()) <= 1e-9 * float(s1_ref.abs().max())
    # and the adjoint the backward kernel uses: d h1 = Mfold^T d s1 (+ the root / gaze path, checked by the GPU gradient tests)
    ds1 = torch.from_numpy(rs.randn(B, 4 * H))
    h1v = h1.clone().requires_grad_(True)
    yv = h1v @ W2.T + b2
    xv = ((yv * os_ + om) - im[:1131]) / is_[:1131]                             # pose channels only (no root / gaze path)
    (xv @ Wx[:, :1131].T * ds1).sum().backward()
    assert float((h1v.grad - ds1 @ Mfold).abs().max()) <= 1e-9 * float(h1v.grad.abs().max())
