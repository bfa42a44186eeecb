#!/usr/bin/env python
"""Fixtures for the JPEG / TSV row (f-4): a small TSV store written the way the reference's tsv_writer does
(src/utils/tsv_file_ops.py:37-53) with base64 JPEGs in the last column, and what the REFERENCE's own reader gives back for
every row -- TSVFile.seek (src/utils/tsv_file.py:64-74) + img_from_base64 (tsv_file_ops.py:21-28) + BGR2RGB
(src/datasets/frei_dataloader.py:203).  Run in the build container (needs /root/reference)."""
import base64
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_bridge as rb  # noqa: E402


def smooth_image(rng, h, w):
    """natural-like content: low-frequency colour field + a few edges + mild noise"""
    base = rng.uniform(0, 255, (5, 5, 3)).astype(np.float32)
    img = cv2.resize(base, (w, h), interpolation=cv2.INTER_CUBIC)
    for _ in range(3):
        x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
        img[y0:y0 + int(rng.integers(4, h // 2)), x0:x0 + int(rng.integers(4, w // 2))] += rng.uniform(-60, 60, 3)
    img += rng.normal(0, 3, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    if not rb.available():
        raise SystemExit("reference tree not mounted")
    rb.load()
    from src.utils.tsv_file import TSVFile as RefTSV
    from src.utils.tsv_file_ops import img_from_base64 as ref_decode
    rng = np.random.default_rng(2025)
    h, w, n = 64, 64, 12
    rows, params = [], []
    for i in range(n):
        rgb = smooth_image(rng, h, w)
        sub = cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444 if i % 2 == 0 else cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420
        ok, enc = cv2.imencode(".jpg", cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR),
                               [cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, sub])
        assert ok
        rows.append([f"img_{i:04d}", '{"k": %d}' % i, base64.b64encode(enc.tobytes()).decode("ascii")])
        params.append(0 if i % 2 == 0 else 1)
    tsv = os.path.join(HERE, "frei_store.tsv")
    with open(tsv, "w") as fp, open(os.path.splitext(tsv)[0] + ".lineidx", "w") as fpidx:
        pos = 0
        for r in rows:                                        # tsv_writer's line format and offsets
            v = "{0}\n".format("\t".join(r))
            fp.write(v)
            fpidx.write(str(pos) + "\n")
            pos += len(v)
    ref = RefTSV(tsv)
    got_rows = [ref.seek(i) for i in range(ref.num_rows())]
    imgs = np.stack([cv2.cvtColor(ref_decode(r[-1]), cv2.COLOR_BGR2RGB) for r in got_rows])
    np.savez_compressed(os.path.join(HERE, "frei_store.npz"), rgb=imgs, subsampling_420=np.array(params, np.uint8),
                        keys=np.array([r[0] for r in got_rows]), col1=np.array([r[1] for r in got_rows]))
    print("rows", len(got_rows), "tsv bytes", os.path.getsize(tsv), "decoded", imgs.shape)


if __name__ == "__main__":
    main()
