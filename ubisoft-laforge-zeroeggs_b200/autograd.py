"""torch.autograd wiring for the C-ABI forward/backward pairs (plumbing only)."""
import torch

from . import _lib
from . import ops

_DEC_GRAD_NAMES = ["dW0", "db0", "dW_ih0", "db_ih0", "dW_hh0", "db_hh0", "dW_ih1", "db_ih1", "dW_hh1", "db_hh1",
                   "dW2", "db2", "dWc0", "dbc0", "dWc1", "dbc1", "dWc2", "dbc2"]


class DecoderWindowFn(torch.autograd.Function):
    """(Y, root_pos, root_rot) = decoder window; backward = zeggs_decoder_window_bwd (full BPTT)."""

    @staticmethod
    def forward(ctx, dec, root_pos0, root_rot0, pose0, gaze_pos, speech, style, in_mean, in_std, out_mean, out_std,
                dt, *weights):
        Y, rp, rq, state = ops.decoder_window_forward(dec, root_pos0, root_rot0, pose0, gaze_pos, speech, style,
                                                      (in_mean, in_std, out_mean, out_std), dt, save=True)
        ctx.dec = dec
        ctx.state = state
        ctx.outs = (Y, rp, rq)
        ctx.weights = weights
        ctx.need_cond = (speech.requires_grad, style.requires_grad)
        return Y, rp, rq

    @staticmethod
    def backward(ctx, dY, dRp, dRq):
        l = _lib.lib()
        a, keep, ws = ctx.state
        dec = ctx.dec
        dev = ctx.outs[0].device
        B, T, H, S, Z = a.B, a.T, a.H, a.S, a.Z
        b = _lib.DecoderBwdArgs()
        hold = []
        for name, g in (("dY", dY), ("dRootPos", dRp), ("dRootRot", dRq)):
            if g is not None:
                g = g.contiguous().float()
                hold.append(g)
                setattr(b, name, g.data_ptr())
        # transposed weight slices for the backward recurrence (cached on the module like the forward pack)
        ver = tuple(p._version for p in dec._weights()) + tuple(p.data_ptr() for p in dec._weights())
        cache = dec.__dict__.get("_zeggs_packed_bwd")
        if cache is None or cache[0] != ver or cache[1].device != dev:
            nb = l.zeggs_decoder_packed_bwd_bytes(H, S, Z)
            packed = torch.empty(nb // 4, dtype=torch.float32, device=dev)
            _lib.check(l.zeggs_decoder_pack_weights_bwd(a, packed.data_ptr(), _lib.stream_ptr()), "zeggs_decoder_pack_weights_bwd")
            dec.__dict__["_zeggs_packed_bwd"] = (ver, packed)
            cache = dec.__dict__["_zeggs_packed_bwd"]
        b.packed_bwd = cache[1].data_ptr()
        grads = [torch.empty_like(w, dtype=torch.float32, memory_format=torch.contiguous_format) for w in ctx.weights]
        for n, g in zip(_DEC_GRAD_NAMES, grads):
            setattr(b, n, g.data_ptr())
        dSpeech = torch.empty((B, T, S), dtype=torch.float32, device=dev)
        dStyle = torch.empty((B, T, Z), dtype=torch.float32, device=dev)
        b.dSpeech, b.dStyle = dSpeech.data_ptr(), dStyle.data_ptr()
        wsb = l.zeggs_decoder_bwd_workspace_bytes(B, T, H, S, Z)
        bws = ops.WS.get("dec_bwd", wsb, dev)
        b.workspace, b.workspace_bytes = bws.data_ptr(), wsb
        _lib.check(l.zeggs_decoder_window_bwd(a, b, _lib.stream_ptr()), "zeggs_decoder_window_bwd")
        ctx.state = None
        return (None, None, None, None, None, dSpeech if ctx.need_cond[0] else None, dStyle if ctx.need_cond[1] else None,
                None, None, None, None, None) + tuple(grads)
