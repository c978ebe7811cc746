"""Development trace of the forward tensor-core recurrence: per-CTA phase durations (SM-local clocks) of a few steps."""
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
_lib.lib().zeggs_debug_set_tc_nacc(int(os.environ.get("VARIANT", "0")))
with torch.no_grad(): dec(*args)
G = H // 8
buf = torch.zeros(G * 64 * 32, dtype=torch.int64, device=dev)
_lib.lib().zeggs_debug_set_tc_trace(buf.data_ptr())
with torch.no_grad(): dec(*args)
torch.cuda.synchronize()
_lib.lib().zeggs_debug_set_tc_trace(None)
tr = buf.cpu().numpy().reshape(G, 64, 32)
names = {0:"L:C(t-1) seen, h1 load issued",4:"M:fold issued",5:"E:d0 ready",6:"E:A done",7:"E:arrived A",8:"L:A seen",10:"M:gh1 issued",
         11:"M:gi0a issued",12:"E:d1 ready",13:"E:B done",14:"L:B seen",16:"M:gi1 issued",17:"E:d2 ready",18:"E:C done",19:"E:arrived C",20:"M:gi0a group0 ready",21:"M:gi0a group1 ready",22:"M:gi0a group2 ready",23:"M:gi0a group3 ready"}
t = 10
base = tr[0, t, 0]
print(f"CTA 0, step {t}: (cycles since 'C(t-1) seen'; 1 us ~ 1900 cyc)")
for ev in sorted(names, key=lambda e: tr[0, t, e]):
    print(f"   {tr[0, t, ev] - base:8d}  {names[ev]}")
print(f"   step period: {tr[0, t+1, 0] - tr[0, t, 0]} cycles")
# per-CTA durations, averaged over steps 8..30
def dur(a, b, nxt=False):
    ts = np.arange(8, 30)
    x = (tr[:, ts + (1 if nxt else 0), b] - tr[:, ts, a]).mean(axis=1)
    return x
rows = [("A: seen -> fold issued", dur(0, 4)), ("A: d0 ready -> epilogue done", dur(5, 6)), ("A: done -> arrived", dur(6, 7)), ("A: arrived -> A seen (wait)", dur(7, 8)),
        ("B: seen -> gi0a issued", dur(8, 11)), ("B: d1 ready -> epilogue done", dur(12, 13)), ("B: done -> B seen (arrive + wait)", dur(13, 14)),
        ("C: seen -> gi1 issued", dur(14, 16)), ("C: d2 ready -> epilogue done", dur(17, 18)), ("C: done -> arrived", dur(18, 19)), ("C: arrived -> next seen (wait)", dur(19, 0, True)),
        ("step period", dur(0, 0, True))]
print(f"{'phase':40s} {'min':>8s} {'median':>8s} {'max':>8s}   argmin argmax (CTA)")
for n, x in rows:
    print(f"{n:40s} {x.min():8.0f} {np.median(x):8.0f} {x.max():8.0f}   {int(x.argmin()):4d} {int(x.argmax()):4d}")
