"""A short train-step run for ncu (launch list / full capture).  Sizes reduced in T so a replayed capture stays short."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from zeggs_b200 import ops
ops.set_decoder_engine(os.environ.get("PROF_ENGINE", "tc"))
T = int(os.environ.get("PROF_T", "32")); steps = int(os.environ.get("PROF_STEPS", "2"))
dev = torch.device("cuda:0")
stepper, P, stats = bench.build_stepper(1024, dev, 1, use_graph=False)      # eager launches: ncu lists every kernel
batch = bench.synth_batch(32, T, int(os.environ.get("PROF_TEX", "384")), seed=1, device=dev)
for _ in range(steps):
    stepper.step(batch)
torch.cuda.synchronize()
print("done")
