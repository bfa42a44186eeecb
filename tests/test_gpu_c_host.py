"""A plain-C program binds include/lighthand_b200.h (no Python, no torch in the process) and runs the host-buffer
pipeline context; its results must equal the CPU oracle's.  This is the boundary a non-Python host uses."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import restate as R

pytestmark = pytest.mark.gpu


def test_c_host_binds_the_abi_and_matches_the_oracle(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_host_demo")
    libdir = os.path.join(ROOT, "lighthand_b200")
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_host_demo.c"), "-o", exe, "-L", libdir,
                    "-llighthand_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True)
    N, S, H, J = 8, 64, 16, 21
    rng = np.random.default_rng(99)
    imgs = rng.integers(0, 256, (N, S, S, 3), dtype=np.uint8)
    joints = rng.uniform(2, 62, (N, J, 2)).astype(np.float32)
    flip = (rng.uniform(size=N) < 0.5)
    M = np.stack([R.gen_trans_from_patch_cv(32 + rng.uniform(-3, 3), 32, 50, 50, S, S, rng.uniform(0.9, 1.1),
                                            rng.uniform(-40, 40)) for _ in range(N)]).astype(np.float64)
    tgt = np.stack([R.generate_target(j, J, (H, H), (4.0, 4.0), 2)[0] for j in joints])
    pred = (tgt + 0.2 * rng.standard_normal(tgt.shape)).astype(np.float32)
    blob = tmp_path / "in.bin"
    with open(blob, "wb") as f:
        for a in (imgs, M, flip.astype(np.uint8), joints, pred):
            f.write(np.ascontiguousarray(a).tobytes())
    out = subprocess.run([exe, str(blob)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])

    loss, _grad = R.joints_mse_loss_and_grad(pred, tgt)
    coords, maxvals, _idx = R.decode_final(pred, 4.0)
    correct, total = R.pck_2d_counts(coords, joints, 0.2, "proportion")
    esum, ecnt = R.epe_train(coords, joints)
    assert got["loss"] == pytest.approx(loss, rel=1e-5)
    assert got["coords_sum"] == pytest.approx(float(coords.astype(np.float64).sum()), rel=1e-9)
    assert got["coord0"] == [float(coords[0, 0, 0]), float(coords[0, 0, 1])]
    assert got["maxval0"] == pytest.approx(float(maxvals[0, 0, 0]), rel=1e-7)
    assert got["counts"] == [3 * (total - correct), 3 * total, 3 * int(ecnt), 3 * N]      # three steps accumulated
    assert got["epe_sum"] == pytest.approx(3 * esum, rel=1e-6)
    assert got["allreduce_ok"] == 1          # lh_comm_* + lh_pipeline_allreduce_metrics at world size 1
    assert got["launches"] == 2 and got["e1"] == -2 and got["e2"] == -2                    # LH_ENULL
