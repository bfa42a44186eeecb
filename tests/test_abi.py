"""The C-ABI library loads and exports exactly what include/lighthand_b200.h declares.  No compute."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from lighthand_b200 import _lib


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "lighthand_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lh_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.PROTOTYPES) == syms, "ctypes prototypes out of sync with the header"


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.lh_abi_version() == _lib.ABI_VERSION == 4
    assert b"NULL" in lib.lh_strerror(-2)
    assert lib.lh_strerror(-1).startswith(b"lighthand_b200")
    assert lib.lh_heatmap_ws_bytes(256, 21) == 16 + 24 * (256 * 21 + 1024) + 16 * 256 * 21


def test_argument_validation_needs_no_device():
    """Argument errors are reported before any CUDA call (negative codes)."""
    lib = _lib.load()
    assert lib.lh_gen_targets(None, 2, 1, 21, 64, 64, 4.0, 4.0, 0, 0, None, 13, None, None, None) == -2
    assert lib.lh_decode(None, 1, 21, 64, 64, 4.0, 0, None, None, None, None, None) == -2
    assert lib.lh_jmse_loss(None, None, None, 1, 21, 64, 64, None, None, None, None, None) == -2
    assert lib.lh_pck_epe_batch(None, None, 2, 1, 21, 0.2, 0, 0, None, None, None) == -2
    assert lib.lh_warp_normalize(None, 0, 1, 1, 3, None, None, None, None, None, 0, 0, 0, None, 1, 1, 1, None) == -2


def test_struct_layout_matches_header():
    # offsets the C side relies on: pointers first, then N (int64), then int32s / floats
    assert ctypes.sizeof(_lib.HeatmapArgs) == 17 * 8 + 8 + 8 * 4 + 4 * 4 + 4 + 4   # 17 pointers, N, 8 int32, 4 float, 2 int32
    assert _lib.HeatmapArgs.N.offset == 136
    assert ctypes.sizeof(_lib.PipelineCfg) == 14 * 4 + 2 * 4 + 6 * 4
    assert ctypes.sizeof(_lib.PipelineIO) == 10 * 8


def test_ops_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from lighthand_b200 import ops
    from lighthand_b200.utils.loss import get_max_preds
    import numpy as np
    with pytest.raises(ops.LighthandError):
        ops.decode(torch.zeros(1, 21, 64, 64))
    with pytest.raises(ops.LighthandError):
        get_max_preds(np.zeros((1, 21, 64, 64), np.float32))


def test_header_is_plain_c_and_the_c_host_links(tmp_path):
    """include/lighthand_b200.h compiles as C11 and the plain-C host (tests/c/abi_host_demo.c, run by the gpu
    suite) links against the library with no C++ or torch symbols needed."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    libdir = os.path.join(ROOT, "lighthand_b200")
    subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_host_demo.c"), "-o", str(tmp_path / "demo"), "-L", libdir,
                    "-llighthand_b200", "-lm"], check=True)


def test_graft_build_entry_point():
    """__graft_entry__.build() is what the driver runs as the does-it-build check: it must agree with the header's
    ABI version (an up-to-date tree makes this a no-op `make`)."""
    import __graft_entry__ as g
    g.build()
