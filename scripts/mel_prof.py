import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
g.build()
from zeggs_b200 import synth, audio
dev = torch.device("cuda:0")
wav = torch.from_numpy(synth.make_waveforms(8, 160000, seed=1)).to(dev).repeat(128, 1)
fe = audio.MelFrontEnd(dev, hop_length=200)
for _ in range(2): fe.forward(wav, 60, 600)
torch.cuda.synchronize()
