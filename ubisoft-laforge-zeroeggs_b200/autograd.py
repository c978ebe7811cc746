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
        ver = ops.weights_key(dec._weights())
        use_tc = a.engine == 1
        if use_tc and l.zeggs_decoder_packed_bwd_tc_bytes(H, S, Z) == 0:
            raise _lib.ZeggsError(f"tensor-core decoder backward unavailable for hidden size {H}")
        if not use_tc:
            cache = dec.__dict__.get("_zeggs_packed_bwd")
            if cache is None or cache[0] != ver or cache[1].device != dev:
                nb = l.zeggs_decoder_packed_bwd_bytes(H, S, Z)
                packed = torch.empty(nb // 4, dtype=torch.float32, device=dev)
                _lib.check(l.zeggs_decoder_pack_weights_bwd(a, packed.data_ptr(), _lib.stream_ptr()), "zeggs_decoder_pack_weights_bwd")
                dec.__dict__["_zeggs_packed_bwd"] = (ver, packed)
                cache = dec.__dict__["_zeggs_packed_bwd"]
            b.packed_bwd = cache[1].data_ptr()
            hold.append(cache[1])
        if use_tc:
            tcc = dec.__dict__.get("_zeggs_packed_bwd_tc")
            if tcc is None or tcc[0] != ver or tcc[1].device != dev:
                ptc = torch.empty(l.zeggs_decoder_packed_bwd_tc_bytes(H, S, Z), dtype=torch.uint8, device=dev)
                _lib.check(l.zeggs_decoder_pack_weights_bwd_tc(a, ptc.data_ptr(), _lib.stream_ptr()), "zeggs_decoder_pack_weights_bwd_tc")
                dec.__dict__["_zeggs_packed_bwd_tc"] = (ver, ptc)
                tcc = dec.__dict__["_zeggs_packed_bwd_tc"]
            wtc = ops.WS.get("dec_bwd_tc", l.zeggs_decoder_bwd_tc_workspace_bytes(H, S, Z), dev)
            b.packed_bwd_tc, b.workspace_tc = tcc[1].data_ptr(), wtc.data_ptr()
            hold += [tcc[1], wtc]
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


class SpeechEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, masks, *weights):
        l = _lib.lib()
        x = ops._f32c(x)
        B, T = x.shape[0], x.shape[1]
        Cin, H, O = weights[0].shape[1], weights[0].shape[0], weights[2].shape[0]
        y = torch.empty((B, T, O), dtype=torch.float32, device=x.device)
        ws = torch.empty(l.zeggs_speech_enc_workspace_bytes(B, T, Cin, H, O), dtype=torch.uint8, device=x.device)
        a, keep = ops.speech_enc_args(enc, x, masks, y, ws)
        _lib.check(l.zeggs_speech_enc_fwd(a, _lib.stream_ptr()), "zeggs_speech_enc_fwd")
        ctx.state = (a, keep, x, masks, y, ws)
        ctx.wshapes = [w.shape for w in weights]
        return y

    @staticmethod
    def backward(ctx, dy):
        a, keep, x, masks, y, ws = ctx.state
        dy = dy.contiguous().float()
        grads = [torch.empty(s, dtype=torch.float32, device=dy.device) for s in ctx.wshapes]
        g = _lib.SpeechEncGrads(dy=dy.data_ptr())
        for n, t in zip(("dW0", "db0", "dW1", "db1", "dW2", "db2"), grads):
            setattr(g, n, t.data_ptr())
        _lib.check(_lib.lib().zeggs_speech_enc_bwd(a, g, _lib.stream_ptr()), "zeggs_speech_enc_bwd")
        ctx.state = None
        return (None, None, None) + tuple(grads)


class StyleEncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, eps, masks, temperature, *weights):
        l = _lib.lib()
        x = ops._f32c(x)
        B, T, Cin = x.shape
        Hs, E = weights[0].shape[0], weights[4].shape[0]
        nh = enc.encoder.blocks[0].attention.multi_head_attention.num_heads
        outs = [torch.empty((B, E // 2), dtype=torch.float32, device=x.device) for _ in range(3)]
        ws = torch.empty(l.zeggs_style_enc_workspace_bytes(B, T, Cin, Hs, E, nh), dtype=torch.uint8, device=x.device)
        a, keep = ops.style_enc_args(enc, x, eps, masks, temperature, outs, ws)
        _lib.check(l.zeggs_style_enc_fwd(a, _lib.stream_ptr()), "zeggs_style_enc_fwd")
        ctx.state = (a, keep, x, eps, masks, outs, ws)
        ctx.wshapes = [w.shape for w in weights]
        return tuple(outs)

    @staticmethod
    def backward(ctx, dz, dmu, dlv):
        a, keep, x, eps, masks, outs, ws = ctx.state
        dev = x.device
        g = _lib.StyleEncGrads()
        hold = []
        for n, t in (("dz", dz), ("dmu", dmu), ("dlogvar", dlv)):
            if t is not None:
                t = t.contiguous().float()
                hold.append(t)
                setattr(g, n, t.data_ptr())
        grads = [torch.empty(s, dtype=torch.float32, device=dev) for s in ctx.wshapes]
        for n, t in zip(_lib.STYLE_W, grads):
            setattr(g, "d" + n, t.data_ptr())
        _lib.check(_lib.lib().zeggs_style_enc_bwd(a, g, _lib.stream_ptr()), "zeggs_style_enc_bwd")
        ctx.state = None
        return (None, None, None, None, None) + tuple(grads)


class TrainLossFn(torch.autograd.Function):
    """loss = train.py:277-421 evaluated (and differentiated) by zeggs_loss_fwd_bwd in the forward call."""

    @staticmethod
    def forward(ctx, Y, rp, rq, WY, Wrp, Wrq, gaze, parents_i32, dt, mu, logvar, kl_weight, terms_out, unit_grad=False, kl_weight_dev=None):
        l = _lib.lib()
        dev = Y.device
        B, T = Y.shape[0], Y.shape[1]
        f = ops._f32c
        Y, rp, rq, WY, Wrp, Wrq, gaze = f(Y), f(rp), f(rq), f(WY), f(Wrp), f(Wrq), f(gaze)
        losses = terms_out if terms_out is not None else torch.empty(19, dtype=torch.float32, device=dev)
        dY, dRp, dRq = torch.empty_like(Y), torch.empty_like(rp), torch.empty_like(rq)
        a = _lib.LossArgs(B=B, T=T, Z=(mu.shape[1] if mu is not None else 0), dt=dt, kl_weight=kl_weight)
        if kl_weight_dev is not None:          # device scalar (graph-replayable): overrides the by-value weight
            a.kl_weight_dev = kl_weight_dev.data_ptr()
        a.Y, a.root_pos, a.root_rot = Y.data_ptr(), rp.data_ptr(), rq.data_ptr()
        a.WY, a.W_root_pos, a.W_root_rot = WY.data_ptr(), Wrp.data_ptr(), Wrq.data_ptr()
        a.gaze_pos, a.parents, a.losses = gaze.data_ptr(), parents_i32.data_ptr(), losses.data_ptr()
        a.dY, a.dRootPos, a.dRootRot = dY.data_ptr(), dRp.data_ptr(), dRq.data_ptr()
        dmu = dlv = None
        if mu is not None:
            mu, logvar = f(mu), f(logvar)
            dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
            a.mu, a.logvar, a.dmu, a.dlogvar = mu.data_ptr(), logvar.data_ptr(), dmu.data_ptr(), dlv.data_ptr()
        wsb = l.zeggs_loss_workspace_bytes(B, T)
        ws = ops.WS.get("loss", wsb, dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
        _lib.check(l.zeggs_loss_fwd_bwd(a, _lib.stream_ptr()), "zeggs_loss_fwd_bwd")
        ctx.grads = (dY, dRp, dRq, dmu, dlv)
        ctx.unit_grad = bool(unit_grad)     # the caller promises loss.backward() with the implicit gradient 1 (TrainStep)
        return losses[0]

    @staticmethod
    def backward(ctx, g):
        dY, dRp, dRq, dmu, dlv = ctx.grads
        ctx.grads = None
        s = (lambda t: t) if ctx.unit_grad else (lambda t: None if t is None else t * g)
        return (s(dY), s(dRp), s(dRq), None, None, None, None, None, None, s(dmu), s(dlv), None, None, None, None)
