"""`modules`-compatible network classes backed by the sm_100a kernels in libzeggs_b200.so.

Drop-in for ZEGGS/modules.py: same class names, constructor arguments, forward signatures and
state-dict keys, so `load_state_dict()` of the shipped checkpoints works and the
reference's `train.py` / `generate.py` can run unchanged with this module registered as
`sys.modules["modules"]`.  The parameter containers are ordinary
`nn.Linear / nn.GRU / nn.Conv1d / nn.LayerNorm / nn.MultiheadAttention` objects (that is what fixes
the key names); their own forward() is never called -- every forward here goes through the C ABI
(csrc/*.cu).  CUDA only: there is no CPU fallback.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import ops

NJ = 75
P_IN = 1134
P_OUT = 1131


# ===============================================================================================
#                                            Decoder
# ===============================================================================================
class RecurrentDecoderNormal(nn.Module):
    """Parameter container of modules.py:165-185 (keys layer0.*, layer1.weight_ih_l0..., layer2.*)."""

    def __init__(self, pose_input_size, speech_size, style_size, output_size, hidden_size, num_rnn_layers):
        super().__init__()
        all_input_size = pose_input_size + speech_size + style_size
        self.layer0 = nn.Linear(all_input_size, hidden_size)
        self.layer1 = nn.GRU(all_input_size + hidden_size, hidden_size, num_rnn_layers, batch_first=True)
        self.layer2 = nn.Linear(hidden_size, output_size)


class CellStateEncoder(nn.Module):
    """Parameter container of modules.py:230-243."""

    def __init__(self, input_size, hidden_size, num_rnn_layers):
        super().__init__()
        self.num_rnn_layers = num_rnn_layers
        self.layer0 = nn.Linear(input_size, hidden_size)
        self.layer1 = nn.Linear(hidden_size, hidden_size)
        self.layer2 = nn.Linear(hidden_size, hidden_size * num_rnn_layers)


class Decoder(nn.Module):
    """modules.py:11-162.  forward() = one launch sequence of zeggs_decoder_window_fwd (+ _bwd under autograd)."""

    def __init__(self, pose_input_size, pose_output_size, speech_encoding_size, style_encoding_size,
                 hidden_size, num_rnn_layers, rnn_cond="normal"):
        super().__init__()
        if rnn_cond != "normal":
            raise _lib.ZeggsError("only rnn_cond='normal' is on the accelerated path (the shipped configs)")
        if num_rnn_layers != 2:
            raise _lib.ZeggsError("num_rnn_layers must be 2 (train.py:124-131 hard-codes it)")
        if pose_input_size != P_IN or pose_output_size != P_OUT:
            raise _lib.ZeggsError("pose layout must be the 75-joint 1134/1131 layout (modules.py:699-736)")
        self.hidden_size = hidden_size
        self.speech_encoding_size = speech_encoding_size
        self.style_encoding_size = style_encoding_size
        self.recurrent_decoder = RecurrentDecoderNormal(
            pose_input_size, speech_encoding_size, style_encoding_size, pose_output_size, hidden_size, num_rnn_layers)
        self.cell_state_encoder = CellStateEncoder(pose_input_size + style_encoding_size, hidden_size, num_rnn_layers)

    def _weights(self):
        r, c = self.recurrent_decoder, self.cell_state_encoder
        return [r.layer0.weight, r.layer0.bias,
                r.layer1.weight_ih_l0, r.layer1.bias_ih_l0, r.layer1.weight_hh_l0, r.layer1.bias_hh_l0,
                r.layer1.weight_ih_l1, r.layer1.bias_ih_l1, r.layer1.weight_hh_l1, r.layer1.bias_hh_l1,
                r.layer2.weight, r.layer2.bias,
                c.layer0.weight, c.layer0.bias, c.layer1.weight, c.layer1.bias, c.layer2.weight, c.layer2.bias]

    def forward(self, Z_root_pos, Z_root_rot, Z_root_vel, Z_root_vrt, Z_lpos, Z_ltxy, Z_lvel, Z_lvrt,
                Z_gaze_pos, speech_encoding, style_encoding, parents, anim_input_mean, anim_input_std,
                anim_output_mean, anim_output_std, dt: float):
        B = speech_encoding.shape[0]
        pose0 = torch.cat([Z_root_vel.reshape(B, -1), Z_root_vrt.reshape(B, -1), Z_lpos.reshape(B, -1),
                           Z_ltxy.reshape(B, -1), Z_lvel.reshape(B, -1), Z_lvrt.reshape(B, -1)], dim=1)
        Y, root_pos, root_rot = ops.decoder_window(
            self, Z_root_pos, Z_root_rot, pose0, Z_gaze_pos, speech_encoding, style_encoding,
            anim_input_mean, anim_input_std, anim_output_mean, anim_output_std, float(dt))
        return ops.split_pose(Y, root_pos, root_rot)


# ===============================================================================================
#                                  Small torch-side helpers kept for API parity
# ===============================================================================================
def normalize(x, eps: float = 1e-8):  # modules.py:672-674
    return x / (torch.norm(x, dim=-1, keepdim=True) + eps)


def generalized_logistic_function(x, center=0.0, B=1.0, A=0.0, K=1.0, C=1.0, Q=1.0, nu=1.0):  # modules.py:745-761
    return A + (K - A) / (C + Q * np.exp(-B * (x - center))) ** (1 / nu)


def compute_KL_div(mu, logvar, iteration):  # modules.py:764-789
    kl_div = torch.mean(-0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp(), dim=1))
    w = min(generalized_logistic_function(iteration, center=7500, B=0.005), 2e-1)
    return kl_div, w
