"""NCCL data parallelism through TrainStep (2 GPUs): gradients after the ONE all-reduce x 1/world == the single-GPU gradients of the
concatenated batch; parameters after the RAdam step identical on every rank.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

from tests._util import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("engine", ["fp32", "tc"])
def test_two_gpu_gradients_equal_single_gpu(engine):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, DP_ENGINE=engine)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_nccl_worker.py")],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("DPRESULT ")][-1]
    d = json.loads(line[len("DPRESULT "):])
    print(" ", d)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"dp_nccl_{engine}.json"), "w") as f:
        json.dump(d, f)
    assert d["ranks_identical"]
    # fp32 engine: summation-order differences only; tc engine: the per-sample bf16 rounding is identical in both runs too
    assert d["grad_rel_l2"] <= (2e-4 if engine == "fp32" else 5e-3)
    assert d["param_max_abs_diff"] <= 1e-6
