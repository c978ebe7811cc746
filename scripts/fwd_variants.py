"""Development probe: forward tensor-core recurrence variants (accumulator counts, per-CTA group rotation) -- device time of one
B=32, T=256, H=1024 window and the deviation of every variant's outputs from the round-1 configuration."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g; g.build()
from zeggs_b200 import ops, synth, _lib
from tests._util import make_decoder, stats_tensors, NAMES, tt
dev = torch.device("cuda:0"); st = stats_tensors()
H, B, T = 1024, 32, int(os.environ.get("T", 256))
P = synth.make_params(H=H, seed=3, with_style=False)
win = tt(synth.make_pose_windows(B, T, seed=3))
gsp = torch.Generator().manual_seed(1)
speech = torch.randn(B, T, 64, generator=gsp) * 0.5; style = torch.randn(B, 1, 64, generator=gsp).repeat(1, T, 1)
dec = make_decoder(P, H, device=dev)
args = [win[n][:, 0].to(dev) for n in NAMES] + [win["gaze_pos"].to(dev), speech.to(dev), style.to(dev), st["parents"]] + \
       [st[k].to(dev) for k in ("anim_input_mean", "anim_input_std", "anim_output_mean", "anim_output_std")] + [st["dt"]]
ops.set_decoder_engine("tc")
lib = _lib.lib()
import ctypes as C
ref = None
res = {}
for v in [int(x) for x in os.environ.get("VARIANTS", "1,0,2").split(",")]:
    lib.zeggs_debug_set_tc_nacc(v)
    with torch.no_grad():
        out = dec(*args)
        torch.cuda.synchronize()
        lib.zeggs_timing_reset(); lib.zeggs_timing_enable(1)
        for _ in range(5):
            out = dec(*args)
        torch.cuda.synchronize()
        lib.zeggs_timing_enable(0)
    tot, cnt = C.c_double(0), C.c_int(0)
    lib.zeggs_timing_read(b"decoder_fwd", C.byref(tot), C.byref(cnt))
    o = [x.float().cpu() for x in out]
    if ref is None:
        ref = o
    dev_ = max(float((a - b).abs().max() / max(1.0, float(b.abs().max()))) for a, b in zip(o, ref))
    res[v] = dict(ms=round(tot.value / max(cnt.value, 1), 4), rel_dev_vs_first=dev_, finite=bool(all(torch.isfinite(x).all() for x in o)))
    print("variant", v, res[v], flush=True)
lib.zeggs_debug_set_tc_nacc(0)
print(json.dumps(res))
