"""Quick device-time probe of the main kernels (CUDA events, warm-up, L2-sized working sets).  Dev tool."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from zeggs_b200 import synth, modules, audio, ops
from tests._util import make_decoder, stats_tensors, NAMES, tt

dev = torch.device("cuda:0")

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

def decoder_case(H, B, T):
    st = stats_tensors(dev)
    P = synth.make_params(H=H, seed=1, with_style=False)
    dec = make_decoder(P, H, device=dev)
    win = tt(synth.make_pose_windows(B, T, seed=1), dev)
    speech = torch.randn(B, T, 64, device=dev) * 0.5
    style = torch.randn(B, 1, 64, device=dev).repeat(1, T, 1)
    args = [win[n][:, 0] for n in NAMES] + [win["gaze_pos"], speech, style, st["parents"], st["anim_input_mean"], st["anim_input_std"], st["anim_output_mean"], st["anim_output_std"], st["dt"]]
    def fwd():
        with torch.no_grad(): dec(*args)
    t_inf = timeit(fwd)
    dec.train()
    sp = speech.clone().requires_grad_(True)
    args2 = list(args); args2[9] = sp
    def fwdbwd():
        out = dec(*args2)
        loss = sum(o.sum() for o in out)
        loss.backward()
    t_fb = timeit(fwdbwd, n=3, warm=1)
    def fwd_save():
        out = dec(*args2)
    t_fs = timeit(fwd_save, n=3, warm=1)
    flop = 2 * (4 * H * 1134 + 12 * H * H + 1131 * H) * B * (T - 1)
    print(f"decoder H={H} B={B} T={T}: infer fwd {t_inf:.2f} ms ({B*(T-1)/t_inf*1e3:.0f} frames/s, {flop/t_inf/1e9:.2f} TFLOP/s, {t_inf/(T-1)*1e3:.1f} us/step) | fwd(save) {t_fs:.2f} ms | fwd+bwd {t_fb:.2f} ms ({B*T/t_fb*1e3:.0f} frames/s)")

for H, B, T in [(1024, 32, 256), (512, 16, 120), (1024, 64, 256), (1024, 1, 600)]:
    decoder_case(H, B, T)

wav = torch.from_numpy(synth.make_waveforms(8, 160000, seed=1)).to(dev).repeat(128, 1)
for hop in (200, 160):
    fe = audio.MelFrontEnd(dev, hop_length=hop)
    t = timeit(lambda: fe.forward(wav, 60, 600))
    L = fe.num_frames(160000)
    byt = wav.numel() * 4 + wav.shape[0] * 600 * 81 * 4
    print(f"mel hop={hop}: {t:.3f} ms for {wav.shape[0]} clips -> {wav.shape[0]/t*1e3:.0f} clips/s, {byt/t/1e6:.0f} GB/s algorithmic (fused 60fps out)")
