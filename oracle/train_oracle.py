"""CPU restatement of one full training step (ZEGGS/train.py:196-432) from the oracle pieces: encoders ->
decoder -> losses -> backward (torch autograd, fp32) -> RAdam.  TEST INFRASTRUCTURE / CPU BASELINE ONLY:
imported by tests/ and by bench.py's cpu_baseline and `--impl reference` legs, never by the product path."""
import time

import numpy as np
import torch

from oracle import model_oracle as mo

NAMES = ["root_pos", "root_rot", "root_vel", "root_vrt", "lpos", "ltxy", "lvel", "lvrt"]


class OracleTrainer:
    def __init__(self, P_np, stats, lr=1e-4, eps=1e-5, label_style=False):
        self.P = {k: torch.from_numpy(np.asarray(v)).clone().requires_grad_(True) for k, v in P_np.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        f = lambda k: torch.as_tensor(stats[k], dtype=torch.float32)
        self.st = {k: f(k) for k in ("audio_input_mean", "audio_input_std", "anim_input_mean", "anim_input_std",
                                     "anim_output_mean", "anim_output_std")}
        self.parents = stats["parents"]
        self.dt = float(stats["dt"])
        self.lr, self.eps = lr, eps
        self.it = 0
        self.label_style = label_style

    def loss(self, batch, eps_vae=None, masks=None):
        s = self.st
        speech = mo.speech_encoder(self.P, (batch["audio"] - s["audio_input_mean"]) / s["audio_input_std"],
                                   None if masks is None else masks.get("speech"))
        mu = logvar = None
        if self.label_style:
            z = batch["style"]
        else:
            z, mu, logvar = mo.style_encoder(self.P, (batch["style"] - s["anim_input_mean"]) / s["anim_input_std"],
                                             eps=eps_vae, masks=None if masks is None else masks.get("style"))
        T = speech.shape[1]
        O = mo.decoder_forward(self.P, *[batch[n][:, 0] for n in NAMES], batch["gaze_pos"], speech,
                               z.unsqueeze(1).repeat(1, T, 1), s["anim_input_mean"], s["anim_input_std"],
                               s["anim_output_mean"], s["anim_output_std"], self.dt)
        loss, terms = mo.train_losses(O, [batch[n] for n in NAMES], batch["gaze_pos"], self.parents, self.dt, mu, logvar, self.it)
        return loss, terms

    def step(self, batch, eps_vae=None, masks=None):
        loss, _ = self.loss(batch, eps_vae, masks)
        keys = [k for k in self.P if not (self.label_style and k.startswith("style_encoder."))]
        grads = torch.autograd.grad(loss, [self.P[k] for k in keys])
        self.it += 1
        with torch.no_grad():
            for k, g in zip(keys, grads):
                mo.radam_step(self.P[k], g, self.m[k], self.v[k], self.it, self.lr, eps=self.eps)
        return float(loss)


def time_cpu_train_step(P_np, stats, batch, min_iters=2, budget_s=20.0, warmup=1, threads=None):
    """frames/s of the CPU port on `batch` (B*T frames per step); returns (frames_per_s, n_iters, threads)."""
    if threads:
        torch.set_num_threads(threads)
    tr = OracleTrainer(P_np, stats)
    for _ in range(warmup):
        tr.step(batch)
    B, T = batch["audio"].shape[0], batch["audio"].shape[1]
    t0 = time.perf_counter()
    n = 0
    while n < min_iters or (time.perf_counter() - t0) < budget_s * 0.5:
        tr.step(batch)
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return B * T * n / dt, n, torch.get_num_threads()
