"""FreiHAND-style TSV image store, read for the GPU.

On-disk format of the reference (``src/utils/tsv_file.py:39-151``, ``tsv_file_ops.py:21-28,37-53``): one sample per
line, tab-separated columns, the LAST column a base64 JPEG (``row[-1]``, src/datasets/frei_dataloader.py:199); a side
file ``<name>.lineidx`` with the byte offset of every line.  ``TSVFile`` keeps the reference's interface (``num_rows``,
``seek``, ``seek_first_column`` / ``get_key``, ``[]``, ``len``) on top of one memory map, and adds what the device path
needs: ``jpeg(idx)`` / ``jpeg_batch(indices)`` hand out the decoded-from-base64 JPEG BYTES without ever building the
Python string columns, and ``decode_batch`` puts the images on the device through ``lh_jpeg_decode_batch`` (nvJPEG) as
``uint8 [N, H, W, 3]`` RGB -- what ``lh_warp_normalize`` / ``lh_frei_rgb_processing`` read.  ``img_from_base64`` stays
available for hosts that want the reference's cv2 image.
"""
from __future__ import annotations

import binascii
import ctypes as C
import mmap
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .._lib import LighthandError, check


def generate_lineidx(filein: str, idxout: str) -> None:
    """Byte offset of every line of ``filein`` -> ``idxout`` (src/utils/tsv_file.py:14-23), one scan of the map."""
    offsets = [0]
    with open(filein, "rb") as f:
        size = os.fstat(f.fileno()).st_size
        if size == 0:
            offsets = []
        else:
            with mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as m:
                pos = m.find(b"\n")
                while pos != -1 and pos + 1 < size:
                    offsets.append(pos + 1)
                    pos = m.find(b"\n", pos + 1)
    tmp = idxout + ".tmp"
    with open(tmp, "w") as out:
        out.write("".join(f"{o}\n" for o in offsets))
    os.rename(tmp, idxout)


def img_from_base64(imagestring):
    """src/utils/tsv_file_ops.py:21-28, verbatim behaviour: base64 -> cv2.imdecode(IMREAD_COLOR) (BGR) or None."""
    import base64
    import cv2
    try:
        jpgbytestring = base64.b64decode(imagestring)
        nparr = np.frombuffer(jpgbytestring, np.uint8)
        return cv2.imdecode(nparr, cv2.IMREAD_COLOR)
    except ValueError:
        return None


class TSVFile(object):
    def __init__(self, tsv_file: str, generate_lineidx: bool = False):
        self.tsv_file = tsv_file
        self.lineidx = os.path.splitext(tsv_file)[0] + ".lineidx"
        self._map = None
        self._fd = None
        self._offsets = None
        self.pid = None
        if not os.path.isfile(self.lineidx) and generate_lineidx:
            globals()["generate_lineidx"](self.tsv_file, self.lineidx)

    def __del__(self):
        self.close()

    def close(self):
        if self._map is not None:
            self._map.close()
            self._map = None
        if self._fd is not None:
            os.close(self._fd)
            self._fd = None

    def __str__(self):
        return "TSVFile(tsv_file='{}')".format(self.tsv_file)

    __repr__ = __str__

    # -- the reference's interface ---------------------------------------------------------------
    def num_rows(self) -> int:
        self._ensure_lineidx_loaded()
        return len(self._offsets)

    def seek(self, idx: int) -> List[str]:
        line = self._line(idx)
        return [s.strip() for s in line.decode("utf-8").split("\t")]

    def seek_first_column(self, idx: int) -> str:
        line = self._line(idx)
        tab = line.find(b"\t")
        return line[: tab if tab != -1 else len(line)].decode("utf-8")

    def get_key(self, idx: int) -> str:
        return self.seek_first_column(idx)

    def __getitem__(self, index: int) -> List[str]:
        return self.seek(index)

    def __len__(self) -> int:
        return self.num_rows()

    # -- what the device path needs -----------------------------------------------------------------
    def jpeg(self, idx: int) -> bytes:
        """The JPEG bit-stream of sample ``idx``: last column, base64-decoded (no intermediate str / list)."""
        line = self._line(idx)
        tab = line.rfind(b"\t")
        return binascii.a2b_base64(line[tab + 1:])

    def jpeg_batch(self, indices: Iterable[int]) -> List[bytes]:
        return [self.jpeg(int(i)) for i in indices]

    # -- internals ----------------------------------------------------------------------------------
    def _line(self, idx: int) -> bytes:
        self._ensure_tsv_opened()
        self._ensure_lineidx_loaded()
        pos = int(self._offsets[idx])
        end = self._map.find(b"\n", pos)
        return self._map[pos: end if end != -1 else len(self._map)]

    def _ensure_lineidx_loaded(self):
        if self._offsets is None:
            with open(self.lineidx, "r") as fp:
                self._offsets = np.array([int(i) for i in fp.read().split()], dtype=np.int64)

    def _ensure_tsv_opened(self):
        if self._map is None or self.pid != os.getpid():     # a forked DataLoader worker opens its own map
            self._fd = os.open(self.tsv_file, os.O_RDONLY)
            self._map = mmap.mmap(self._fd, 0, access=mmap.ACCESS_READ)
            self.pid = os.getpid()


class JpegDecoder:
    """Batched JPEG decode on the device through the C ABI (``lh_jpeg_*``, nvJPEG underneath)."""

    BACKENDS = {"default": 0, "hybrid": 1, "gpu_hybrid": 2, "hardware": 3}

    def __init__(self, device: Optional[int] = None, backend: str = "gpu_hybrid", cpu_threads: int = 1):
        if not torch.cuda.is_available():
            raise LighthandError("JpegDecoder needs a CUDA device (there is no CPU fallback in lighthand_b200)")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.backend = backend
        self._h = C.c_void_p()
        check(self.lib.lh_jpeg_create(self.device, self.BACKENDS[backend], int(cpu_threads), C.byref(self._h)),
              f"lh_jpeg_create({backend})")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.lh_jpeg_destroy(self._h)
            self._h = None

    __del__ = close

    def decode(self, streams: Sequence[bytes], hw, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``streams``: N JPEG bit-streams of size hw = (H, W) -> uint8 [N, H, W, 3] RGB on the device (asynchronous on
        torch's current stream; the byte strings are kept alive until the next call or ``close``)."""
        n, (h, w) = len(streams), hw
        if out is None:
            out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=f"cuda:{self.device}")
        if tuple(out.shape) != (n, h, w, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
            raise LighthandError("JpegDecoder.decode: out must be a contiguous uint8 [N, H, W, 3] CUDA tensor")
        ptrs = (C.c_char_p * n)(*streams)
        sizes = (C.c_size_t * n)(*[len(s) for s in streams])
        self._keep = (list(streams), ptrs, sizes)
        st = torch.cuda.current_stream(self.device).cuda_stream
        check(self.lib.lh_jpeg_decode_batch(self._h, ptrs, sizes, n, w, h, out.data_ptr(), h * w * 3, w * 3, st),
              "lh_jpeg_decode_batch")
        return out


def decode_batch(tsv: TSVFile, indices: Iterable[int], hw, decoder: Optional[JpegDecoder] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Rows ``indices`` of the store as one uint8 [N, H, W, 3] RGB device tensor = ``get_image`` of the reference
    (src/datasets/frei_dataloader.py:195-204, RGB branch) for every row, stacked, without a host-side decode."""
    dec = decoder or JpegDecoder()
    return dec.decode(tsv.jpeg_batch(indices), hw, out)
