"""Worker of tests/test_dp_nccl_gpu.py (launched by torch.distributed.run, one rank per GPU): data-parallel gradients through
TrainStep == single-GPU gradients on the concatenated batch."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", lrank)
    torch.cuda.set_device(dev)
    torch.distributed.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    torch.distributed.barrier()
    from zeggs_b200 import dp, ops
    from bench import build_stepper, synth_batch
    engine = os.environ.get("DP_ENGINE", "fp32")
    ops.set_decoder_engine(engine)
    H, Bg, T, T_ex = 384, 8, 10, 16
    full = synth_batch(Bg, T, T_ex, seed=7)
    eps = torch.from_numpy(np.random.RandomState(1).randn(Bg, 64).astype(np.float32))
    lo, hi = dp.shard_range(Bg, rank, world)
    # data parallel: each rank its shard, one all-reduce, 1/world folded into the optimizer
    st_dp, _, _ = build_stepper(H, dev, world, use_graph=False)
    st_dp.iteration = 9000
    shard = {k: v[lo:hi].to(dev) for k, v in full.items()}
    st_dp.optimizer.zero_grad()
    st_dp.forward_backward(shard, eps=eps[lo:hi].to(dev), train_mode=False)
    st_dp._allreduce()
    g_dp = (st_dp.optimizer.flat_grad * st_dp.optimizer.grad_scale).clone()
    st_dp.optimizer.step()
    p_dp = st_dp.optimizer.flat_param.clone()
    # single GPU, whole batch
    st_1, _, _ = build_stepper(H, dev, 1, use_graph=False)
    st_1.iteration = 9000
    st_1.optimizer.zero_grad()
    st_1.forward_backward({k: v.to(dev) for k, v in full.items()}, eps=eps.to(dev), train_mode=False)
    g_1 = st_1.optimizer.flat_grad.clone()
    st_1.optimizer.step()
    p_1 = st_1.optimizer.flat_param.clone()
    torch.cuda.synchronize()
    rel_g = float((g_dp - g_1).norm() / g_1.norm())
    max_g = float((g_dp - g_1).abs().max() / g_1.abs().max())
    rel_p = float((p_dp - p_1).abs().max())
    # every rank holds the same parameters after the step
    chk = p_dp.double().sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    torch.distributed.all_gather(allc, chk)
    same = all(float(c) == float(allc[0]) for c in allc)
    if rank == 0:
        print("DPRESULT " + json.dumps(dict(world=world, engine=engine, grad_rel_l2=rel_g, grad_max_rel=max_g, param_max_abs_diff=rel_p,
                                            ranks_identical=same)))
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
