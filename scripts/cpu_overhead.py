"""Host-side enqueue time of one train step (no device synchronisation inside the loop).  Dev tool."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from zeggs_b200 import ops
ops.set_decoder_engine("tc")
dev = torch.device("cuda:0")
stepper, P, stats = bench.build_stepper(1024, dev, 1)
batch = bench.synth_batch(32, 256, 384, seed=1, device=dev)
for _ in range(5): stepper.step(batch)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10): stepper.step(batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step, load {os.getloadavg()}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): stepper.step(batch)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
