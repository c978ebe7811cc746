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
_lib.lib().zeggs_debug_set_tc_nacc(int(os.environ.get("NACC", "0")))
with torch.no_grad(): dec(*args)
buf = torch.zeros(64 * 32, dtype=torch.int64, device=dev)
_lib.lib().zeggs_debug_set_tc_trace(buf.data_ptr())
with torch.no_grad(): dec(*args)
torch.cuda.synchronize()
_lib.lib().zeggs_debug_set_tc_trace(None)
tr = buf.cpu().numpy().reshape(64, 32)
names = {0:"L:B4 seen",1:"L:xp load issued",2:"M:before xa wait",3:"M:xa(S1) ready",4:"M:S1 issued",5:"E:d0 ready",6:"E:s1 done",7:"E:arrived B1",8:"L:B1 seen",
         9:"M:before xa wait",10:"M:xa(a) ready",11:"M:gi0a issued",12:"E:d1 ready",13:"E:s2 done",14:"L:B2 seen",15:"M:xa(h0) ready",16:"M:gi1 issued",17:"E:d2 ready",
         18:"E:s3 done",19:"E:arrived B3",20:"L:B3 seen",21:"M:xa(h1) ready",22:"E:d3 ready",23:"E:s4 done",24:"E:arrived B4"}
for t in (10, 20):
    base = tr[t, 0]
    print(f"step {t}: (cycles since 'B4 seen'; 1 us ~ 1900 cyc)")
    for ev in sorted(names, key=lambda e: tr[t, e]) if t == 10 else []:
        print(f"   {tr[t, ev] - base:8d}  {names[ev]}")
    print(f"   step period: {tr[t+1,0]-tr[t,0]} cycles")
