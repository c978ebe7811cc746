"""RAdam with the reference's constructor surface (ZEGGS/optimizers.py) backed by one fused kernel.

All parameters are re-homed into ONE flat fp32 buffer (each nn.Parameter becomes a view of it), with a
matching flat gradient buffer: the optimizer step is a single launch of zeggs_radam_step and the
data-parallel gradient exchange is a single NCCL all-reduce of `flat_grad`."""
import torch

from . import _lib
from . import ops


class RAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=True):
        if weight_decay != 0 or not degenerated_to_sgd:
            raise _lib.ZeggsError("fused RAdam implements weight_decay=0, degenerated_to_sgd=True (train.py:160)")
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0))
        ps = [p for g in self.param_groups for p in g["params"]]
        dev = ps[0].device
        if dev.type != "cuda":
            raise _lib.ZeggsError("fused RAdam needs CUDA parameters (no CPU fallback)")
        n = sum(p.numel() for p in ps)
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in ps:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view(p.shape)
            p.grad = self.flat_grad[off:off + k].view(p.shape)
            off += k
        self._step = 0
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self._step += 1
        _lib.check(_lib.lib().zeggs_radam_step(
            self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.flat_param.numel(), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
            int(self._step), float(self.grad_scale), _lib.stream_ptr()), "zeggs_radam_step")
        ops.bump_weights_epoch()     # parameters changed behind autograd's back: invalidate packed-weight caches
        return loss

    def state_dict(self):
        d = super().state_dict()
        d["zeggs"] = dict(step=self._step, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq)
        return d

    def load_state_dict(self, state_dict):
        z = state_dict.get("zeggs")
        if z is not None:
            self._step = int(z["step"])
            self.exp_avg.copy_(z["exp_avg"])
            self.exp_avg_sq.copy_(z["exp_avg_sq"])
            return
        # reference checkpoint (per-parameter state, optimizers.py:47-56)
        ps = [p for g in self.param_groups for p in g["params"]]
        off = 0
        for i, p in enumerate(ps):
            st = state_dict["state"].get(i)
            k = p.numel()
            if st is not None:
                self._step = int(st["step"])
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
            off += k
