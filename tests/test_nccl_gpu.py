"""Multi-GPU NCCL correctness, run where at least two GPUs are visible (`gpurun --gpus 2`); skipped on
a single-GPU box.  The check itself is tools/nccl_check.py, one process per GPU under torch.distributed.run."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
def test_sharded_inference_and_gradient_allreduce_over_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "nccl_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=540)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["inference_equal"], res
    assert res["ranks_differ_before"] and res["allreduce_mean_rel_err"] < 1e-6, res
    assert res["weights_identical_after_steps"] and res["weights_finite"] and res["ok"], res
