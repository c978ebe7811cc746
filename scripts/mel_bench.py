"""BASELINE.json config 3: mel-spectrogram microbench, 1024 x 16 kHz 10 s clips, 80 mels (device time, CUDA events).  Dev tool."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from zeggs_b200 import synth, audio
dev = torch.device("cuda:0")
wav = torch.from_numpy(synth.make_waveforms(8, 160000, seed=1)).to(dev).repeat(128, 1)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
out = {}
for hop in (200, 160):
    fe = audio.MelFrontEnd(dev, hop_length=hop)
    L = fe.num_frames(160000)
    for mode in ("feat60", "mel"):
        fn = (lambda: fe.forward(wav, 60, 600)) if mode == "feat60" else (lambda: fe.forward(wav, want_mel=True, want_feat=False))
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = float(np.median(ts))
        byt = wav.numel() * 4 + (wav.shape[0] * 600 * 81 * 4 if mode == "feat60" else wav.shape[0] * 80 * L * 4)
        out[f"hop{hop}_{mode}"] = dict(ms=round(t, 3), clips_per_s=round(wav.shape[0] / t * 1e3), algorithmic_gbs=round(byt / t / 1e6, 1),
                                       frac_of_hbm=round(byt / t / 1e6 / peaks["hbm_gbs"], 4), frames=L)
print(json.dumps(out))
