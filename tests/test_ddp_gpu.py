"""GPU, world_size 2 over NCCL (skipped on a 1-GPU box): N-rank batch-sharded DDP == 1-rank on the concatenated batch, eagerly
(tools/ddp_check.py, incl. finish_reduce() + gradient clipping) and with the whole step captured in a CUDA graph
(tools/ddp_graph_check.py).  SURVEY.md §4: "N-rank DDP gradients == 1-rank gradients on the concatenated batch"."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc=2, timeout=600):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("script,ok", [("ddp_check.py", "DDP CHECK OK"), ("ddp_graph_check.py", "OK")])
def test_two_rank_ddp_matches_single_process(script, ok):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun(script)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert ok in r.stdout
