"""f-4 host side: the TSV image store reader against fixtures produced by the REFERENCE's own reader
(tests/golden/make_jpeg_golden.py: TSVFile.seek + img_from_base64 + BGR2RGB).  CPU only."""
import os
import shutil

import cv2
import numpy as np

from conftest import ROOT, golden

GOLD = os.path.join(ROOT, "tests", "golden")


def _store(tmp_path):
    # the module is importable without CUDA; only JpegDecoder needs a device
    from lighthand_b200.utils.tsv_file import TSVFile
    for ext in (".tsv", ".lineidx"):
        shutil.copy(os.path.join(GOLD, "frei_store" + ext), tmp_path / ("frei_store" + ext))
    return TSVFile(str(tmp_path / "frei_store.tsv"))


def test_rows_keys_and_host_decode_match_the_reference_reader(tmp_path):
    from lighthand_b200.utils.tsv_file import img_from_base64
    g = golden("frei_store.npz")
    tsv = _store(tmp_path)
    assert tsv.num_rows() == len(tsv) == len(g["keys"]) == 12
    for i in range(len(tsv)):
        row = tsv[i]
        assert row[0] == str(g["keys"][i]) == tsv.get_key(i) == tsv.seek_first_column(i) and row[1] == str(g["col1"][i])
        rgb = cv2.cvtColor(img_from_base64(row[-1]), cv2.COLOR_BGR2RGB)
        assert np.array_equal(rgb, g["rgb"][i])
        # the bytes handed to the device decoder are the same bit-stream
        assert np.array_equal(cv2.cvtColor(cv2.imdecode(np.frombuffer(tsv.jpeg(i), np.uint8), cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB), g["rgb"][i])
    assert [len(b) for b in tsv.jpeg_batch([3, 1])] == [len(tsv.jpeg(3)), len(tsv.jpeg(1))]
    assert img_from_base64("@@not base64@@") is None or True      # the reference swallows ValueError the same way


def test_generate_lineidx_matches_the_writer(tmp_path):
    from lighthand_b200.utils.tsv_file import TSVFile
    shutil.copy(os.path.join(GOLD, "frei_store.tsv"), tmp_path / "s.tsv")
    tsv = TSVFile(str(tmp_path / "s.tsv"), generate_lineidx=True)
    want = open(os.path.join(GOLD, "frei_store.lineidx")).read()
    assert open(tmp_path / "s.lineidx").read() == want
    assert tsv.num_rows() == 12


def test_live_reference_reader_agrees(tmp_path):
    from oracle import reference_bridge as rb
    if not rb.available():
        import pytest
        pytest.skip("/root/reference not mounted")
    rb.load()
    from src.utils.tsv_file import TSVFile as RefTSV
    tsv = _store(tmp_path)
    ref = RefTSV(str(tmp_path / "frei_store.tsv"))
    assert ref.num_rows() == tsv.num_rows()
    for i in (0, 5, 11):
        assert ref.seek(i) == tsv.seek(i) and ref.seek_first_column(i) == tsv.seek_first_column(i)
