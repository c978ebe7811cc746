"""BASELINE.json config 5: batch inference, 64 concurrent 60 s clips (T = 3600), autoregressive decoder (device time).  Dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from zeggs_b200 import synth, ops
from tests._util import make_decoder, stats_tensors, NAMES, tt
dev = torch.device("cuda:0"); st = stats_tensors(dev)
B, T, H = int(os.environ.get("B", 64)), int(os.environ.get("T", 3600)), 1024
P = synth.make_params(H=H, seed=1, with_style=False)
dec = make_decoder(P, H, device=dev)
win = tt(synth.make_pose_windows(B, 2, seed=1), dev)
gaze = torch.from_numpy(synth.make_pose_windows(B, T, seed=2)["gaze_pos"]).to(dev)
speech = torch.randn(B, T, 64, device=dev) * 0.5
style = torch.randn(B, 1, 64, device=dev).repeat(1, T, 1)
args = [win[n][:, 0] for n in NAMES] + [gaze, speech, style, st["parents"], st["anim_input_mean"], st["anim_input_std"], st["anim_output_mean"], st["anim_output_std"], st["dt"]]
out = {}
for eng in ("tc", "fp32"):
    ops.set_decoder_engine(eng)
    ts = []
    with torch.no_grad():
        for i in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o = dec(*args); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts[1:]))
    out[eng] = dict(ms=round(t, 2), frames_per_s=round(B * (T - 1) / t * 1e3), finite=bool(all(torch.isfinite(x).all() for x in o)))
print(json.dumps(dict(config="batch inference B=%d T=%d H=%d" % (B, T, H), **out)))
