"""GPU, >= 2 devices: data-parallel replicas under NCCL stay BIT-IDENTICAL and the bucketed, overlapped gradient exchange
of the prefetch schedule follows the parameter trajectory of the sequential schedule with one all-reduce after the
backward (trainer/leo_trainer.py:50-52,135: DDP around the trainable prompter).  Launches tools/check_ddp_equiv.py under
torch.distributed.run with two ranks; skipped on a one-GPU box (the gloo tests in tests/test_dist_cpu.py cover the
host-side logic there)."""
import os
import subprocess
import sys

import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_replicas_identical_and_bucketed_exchange_equals_single_allreduce():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(helpers.ROOT, "tools", "check_ddp_equiv.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=helpers.ROOT, timeout=600)
    assert r.returncode == 0 and "ddp equivalence ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
