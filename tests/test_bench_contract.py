"""bench.py's reference arm runs on the CPU and prints the contract's JSON line (no GPU needed)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_contract_line():
    env = {**os.environ, "OMP_NUM_THREADS": "1"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "heatmap_pipeline_images_per_s" and line["unit"] == "images/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    # the reference's own functions when its tree is reachable (live mount here, baseline/_ref on the GPU box), else the port
    from oracle import reference_arm
    assert line["cpu_baseline"]["kind"] == ("reference" if reference_arm.locate() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("C2")


def test_reference_arm_other_ranks_exit_quietly():
    env = {**os.environ, "RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_port_pass_and_reference_pass_do_the_same_work():
    """The two CPU arms (reference functions / oracle port) run the same stages on the same batch without error; the
    port is what the arm falls back to when baseline/_ref is absent."""
    import numpy as np
    import torch
    import bench
    from oracle import reference_arm, restate as R
    d = bench.synth_host(4, 256, 3)
    pred = np.random.default_rng(0).standard_normal((4, 21, 64, 64)).astype(np.float32)
    reference_arm.make_pass(None, R, torch)((*d, pred), 256, 64, 2, True)
    ns, kind = reference_arm.load()
    if ns is not None:
        reference_arm.make_pass(ns, R, torch)((*d, pred), 256, 64, 2, True)
        reference_arm.make_pass(ns, R, torch)((*d, pred[:, :, :56, :56].copy()), 224, 56, 2, False)
