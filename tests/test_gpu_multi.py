"""Multi-GPU parity on hardware: an N-rank run of the path, all-reduced with the C ABI's NCCL collective, equals the
1-rank run on the identical global data -- bit for bit for the integer counters and curve histograms, 1e-12 for the
float64 sums.  Needs >= 2 GPUs (``gpurun --gpus 2`` / 4 / 8); skipped on a single-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_equals_single_rank_run(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    port = 29600 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    got = json.loads(line)
    assert got["ok"] and got["world"] == world and got["counts"][3] == 1000
