"""f-4 device side: batched JPEG decode into K1's source layout (lh_jpeg_decode_batch) against what the reference's
reader returns (cv2.imdecode via img_from_base64, fixtures in tests/golden/frei_store.*).

Parity bar: a JPEG decoder is specified only up to IDCT and chroma-upsampling accuracy, so nvJPEG and libjpeg-turbo are
NOT bit-identical.  Stated bar: 4:4:4 streams max |diff| <= 3 LSB and mean |diff| <= 0.6; 4:2:0 streams
(chroma interpolated on upsampling in both decoders, different filters) max |diff| <= 5 LSB and mean |diff| <= 0.9.
Measured on B200 / nvJPEG 12.4 against OpenCV 4.13 (libjpeg-turbo): 3 / 0.50 and 4 / 0.69."""
import os
import shutil

import numpy as np
import pytest
import torch

from conftest import ROOT, golden

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("backend", ["gpu_hybrid", "hybrid", "hardware"])
def test_decode_batch_vs_reference_reader(tmp_path, backend):
    from lighthand_b200.utils.tsv_file import JpegDecoder, TSVFile, decode_batch
    from lighthand_b200._lib import LighthandError
    for ext in (".tsv", ".lineidx"):
        shutil.copy(os.path.join(GOLD, "frei_store" + ext), tmp_path / ("frei_store" + ext))
    tsv = TSVFile(str(tmp_path / "frei_store.tsv"))
    g = golden("frei_store.npz")
    try:
        dec = JpegDecoder(backend=backend)
    except LighthandError:
        if backend == "hardware":
            pytest.skip("NVJPG hardware backend not available on this part / driver")
        raise
    idx = list(range(len(tsv)))
    try:
        out = decode_batch(tsv, idx, (64, 64), dec)
        torch.cuda.synchronize()
    except LighthandError:
        if backend == "hardware":
            pytest.skip("NVJPG hardware backend rejected the batch")
        raise
    got = out.cpu().numpy().astype(np.int32)
    want = g["rgb"].astype(np.int32)
    assert got.shape == want.shape
    d = np.abs(got - want)
    s420 = g["subsampling_420"].astype(bool)
    d444, d420 = d[~s420], d[s420]
    print(f"[{backend}] 4:4:4 max {d444.max()} mean {d444.mean():.3f} | 4:2:0 max {d420.max()} mean {d420.mean():.3f} "
          f"p99.9 {np.percentile(d420, 99.9):.1f}")
    assert d444.max() <= 3 and d444.mean() <= 0.6
    assert d420.max() <= 5 and d420.mean() <= 0.9
    # a second batch of another size re-initialises the batched decoder; results unchanged
    out2 = decode_batch(tsv, [4, 2, 9], (64, 64), dec)
    assert torch.equal(out2, out[[4, 2, 9]])
    dec.close()


def test_threaded_hybrid_equals_single_threaded(tmp_path):
    """cpu_threads > 1 spreads the host Huffman stage of the HYBRID backend over worker threads (decoupled nvJPEG API, one
    state per thread); same decoder, so the pixels must not change -- also with fewer images than threads and across batches."""
    from lighthand_b200.utils.tsv_file import JpegDecoder, TSVFile, decode_batch
    for ext in (".tsv", ".lineidx"):
        shutil.copy(os.path.join(GOLD, "frei_store" + ext), tmp_path / ("frei_store" + ext))
    tsv = TSVFile(str(tmp_path / "frei_store.tsv"))
    one = JpegDecoder(backend="hybrid", cpu_threads=1)
    many = JpegDecoder(backend="hybrid", cpu_threads=6)
    idx = list(range(len(tsv))) * 3
    want = decode_batch(tsv, idx, (64, 64), one)
    for _ in range(3):
        got = decode_batch(tsv, idx, (64, 64), many)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    assert torch.equal(decode_batch(tsv, [5, 1], (64, 64), many), want[[5, 1]])
    one.close()
    many.close()


def test_decoded_batch_feeds_k1_directly(tmp_path):
    """decode -> lh_warp_normalize with no host round trip: identity crop of the decoded images == ToTensor + Normalize."""
    from lighthand_b200 import ops
    from lighthand_b200.utils.tsv_file import JpegDecoder, TSVFile, decode_batch
    for ext in (".tsv", ".lineidx"):
        shutil.copy(os.path.join(GOLD, "frei_store" + ext), tmp_path / ("frei_store" + ext))
    tsv = TSVFile(str(tmp_path / "frei_store.tsv"))
    imgs = decode_batch(tsv, range(12), (64, 64), JpegDecoder())
    M = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]], dtype=torch.float64, device="cuda").repeat(12, 1, 1)
    out = ops.warp_normalize(imgs, M, (64, 64), exact_norm=True)
    mean = torch.tensor(ops.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGENET_STD).view(1, 3, 1, 1)
    want = (imgs.cpu().permute(0, 3, 1, 2).float() / 255 - mean) / std      # ToTensor + Normalize as torch's CPU ops round
    assert torch.equal(out.cpu(), want)


def test_wrong_size_is_reported(tmp_path):
    from lighthand_b200.utils.tsv_file import JpegDecoder, TSVFile
    from lighthand_b200._lib import LighthandError
    for ext in (".tsv", ".lineidx"):
        shutil.copy(os.path.join(GOLD, "frei_store" + ext), tmp_path / ("frei_store" + ext))
    tsv = TSVFile(str(tmp_path / "frei_store.tsv"))
    with pytest.raises(LighthandError):
        JpegDecoder().decode(tsv.jpeg_batch([0]), (32, 32))
    threaded = JpegDecoder(backend="hybrid", cpu_threads=4)      # the threaded mode checks inside its workers
    with pytest.raises(LighthandError):
        threaded.decode(tsv.jpeg_batch([0, 1, 2]), (32, 32))
    assert threaded.decode(tsv.jpeg_batch([0, 1, 2]), (64, 64)).shape == (3, 64, 64, 3)      # and stays usable
    threaded.close()
