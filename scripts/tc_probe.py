import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g; g.build()
from zeggs_b200 import ops, synth
from tests._util import make_decoder, stats_tensors, NAMES, tt
from oracle import model_oracle as mo
dev = torch.device("cuda:0")
st = stats_tensors()
def case(H, B, T, seed=3, time_it=False):
    P = synth.make_params(H=H, seed=seed, with_style=False)
    win = tt(synth.make_pose_windows(B, T, seed=seed))
    rs = np.random.RandomState(seed)
    speech = torch.from_numpy((rs.randn(B, T, 64) * 0.5).astype(np.float32))
    style = torch.from_numpy(rs.randn(B, 1, 64).astype(np.float32)).repeat(1, T, 1)
    dec = make_decoder(P, H, device=dev)
    args = [win[n][:, 0].to(dev) for n in NAMES] + [win["gaze_pos"].to(dev), speech.to(dev), style.to(dev), st["parents"]] + \
           [st[k].to(dev) for k in ("anim_input_mean", "anim_input_std", "anim_output_mean", "anim_output_std")] + [st["dt"]]
    outs = {}
    for eng in ("fp32", "tc"):
        ops.set_decoder_engine(eng)
        with torch.no_grad():
            outs[eng] = [o.clone() for o in dec(*args)]
        torch.cuda.synchronize()
        if time_it:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.no_grad():
                dec(*args); e0.record(); dec(*args); dec(*args); e1.record()
            torch.cuda.synchronize()
            print(f"   {eng}: {e0.elapsed_time(e1)/2:.3f} ms per window ({e0.elapsed_time(e1)/2/(T-1)*1e3:.1f} us/step)")
    sd = st["anim_output_std"]
    for i, n in enumerate(NAMES):
        a, b = outs["tc"][i].cpu(), outs["fp32"][i].cpu()
        d = (a - b).abs()
        print(f"  H{H} B{B} T{T} {n:9s} max|tc-fp32| {d.max().item():.3e} (ref max {b.abs().max().item():.3e}) nan={torch.isnan(a).sum().item()}  last-frame err {d[:, -1].max().item():.3e}")
for cfg in [(128, 4, 9), (512, 16, 12), (1024, 32, 8)]:
    case(*cfg)
case(1024, 32, 128, time_it=True)
case(512, 16, 120, time_it=True)
