import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g; g.build()
from zeggs_b200 import ops, synth, _lib
from tests._util import make_decoder, stats_tensors, NAMES, tt
dev = torch.device("cuda:0"); st = stats_tensors()
H, B, T = int(os.environ.get("H", 1024)), int(os.environ.get("B", 32)), 40
P = synth.make_params(H=H, seed=3, with_style=False)
win = tt(synth.make_pose_windows(B, T, seed=3))
speech = torch.randn(B, T, 64) * 0.5; style = torch.randn(B, 1, 64).repeat(1, T, 1)
dec = make_decoder(P, H, device=dev)
args = [win[n][:, 0].to(dev) for n in NAMES] + [win["gaze_pos"].to(dev), speech.to(dev), style.to(dev), st["parents"]] + \
       [st[k].to(dev) for k in ("anim_input_mean", "anim_input_std", "anim_output_mean", "anim_output_std")] + [st["dt"]]
ops.set_decoder_engine("tc")
_lib.lib().zeggs_debug_set_tc_cluster(int(os.environ.get("CLUSTER", "8")))
with torch.no_grad(): dec(*args)
buf = torch.zeros(64 * 32, dtype=torch.int64, device=dev)
_lib.lib().zeggs_debug_set_tc_trace(buf.data_ptr())
with torch.no_grad(): dec(*args)
torch.cuda.synchronize()
_lib.lib().zeggs_debug_set_tc_trace(None)
tr = buf.cpu().numpy().reshape(64, 32)
print("cluster size used:", _lib.lib().zeggs_debug_get_tc_cluster())
names = {0:"L:C(t-1) seen, h1 load issued",4:"M:fold issued",5:"E:d0 ready",6:"E:A done",7:"E:arrived A",8:"L:A seen",10:"M:gh1 issued",
         11:"M:gi0a issued",12:"E:d1 ready",13:"E:B done",14:"L:B seen",16:"M:gi1 issued",17:"E:d2 ready",18:"E:C done",19:"E:arrived C",20:"M:gi0a group0 ready",21:"M:gi0a group1 ready",22:"M:gi0a group2 ready",23:"M:gi0a group3 ready"}
for t in (10, 20):
    base = tr[t, 0]
    print(f"step {t}: (cycles since 'B4 seen'; 1 us ~ 1900 cyc)")
    for ev in sorted(names, key=lambda e: tr[t, e]) if t == 10 else []:
        print(f"   {tr[t, ev] - base:8d}  {names[ev]}")
    print(f"   step period: {tr[t+1,0]-tr[t,0]} cycles")
