"""Optional device-side JPEG codec around the warp (nvJPEG through the C ABI, include/vr180_b200.h section 6).

The reference reads and writes image files with OpenCV on the host (remapper.py:373, :453, :519).  With
`set_codec("nvjpeg")` the JPEG -> warp -> JPEG form of `apply_lr` keeps every uncompressed frame on the GPU: the file
bytes go up, nvJPEG decodes them into device memory, the warp kernel writes the SBS frame next to them, nvJPEG encodes
it there and only the compressed stream comes back.  OPT-IN: nvJPEG's decoder differs from the libjpeg-turbo decoder
inside cv.imread by a few grey levels on some pixels, so files that go through it are not bit-identical to the
reference's; the default codec stays cv2.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Any

import numpy as np

from . import _native as N

_JPEG_SUFFIXES = (".jpg", ".jpeg", ".jpe")


def available() -> bool:
    return bool(N.lib().vr180_jpeg_available())


def is_jpeg_path(p: Any) -> bool:
    return isinstance(p, (str, Path)) and Path(p).suffix.lower() in _JPEG_SUFFIXES


def decode_jpeg_device(data: bytes | str | Path, device: Any = None):
    """JPEG file / bytes -> (H, W, 3) uint8 CUDA tensor, interleaved BGR like cv.imread, rows 16-byte aligned."""
    import torch

    if isinstance(data, (str, Path)):
        data = Path(data).read_bytes()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    buf = np.frombuffer(data, dtype=np.uint8)
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    with torch.cuda.device(dev):
        N.check(N.lib().vr180_jpeg_info(buf.ctypes.data, buf.size, C.byref(w), C.byref(h), C.byref(ch)), "vr180_jpeg_info")
        pitch = (w.value * 3 + 15) // 16 * 16
        store = torch.empty((h.value, pitch), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        N.check(N.lib().vr180_jpeg_decode(buf.ctypes.data, buf.size, store.data_ptr(), pitch, w.value, h.value, stream),
                "vr180_jpeg_decode")
    return store.as_strided((h.value, w.value, 3), (pitch, 3, 1))


def encode_jpeg_device(image, quality: int = 95) -> bytes:
    """(H, W, 3) uint8 CUDA tensor (BGR, pixel-contiguous rows) -> JPEG bytes with cv.imwrite's defaults (4:2:0)."""
    import torch

    if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3 or image.stride(2) != 1 or image.stride(1) != 3:
        raise ValueError("encode_jpeg_device takes an (H, W, 3) uint8 CUDA tensor with contiguous pixels")
    h, w = int(image.shape[0]), int(image.shape[1])
    out = np.empty(w * h * 3 // 2 + (1 << 16), dtype=np.uint8)
    n = C.c_size_t(out.size)
    with torch.cuda.device(image.device):
        stream = torch.cuda.current_stream(image.device).cuda_stream
        N.check(N.lib().vr180_jpeg_encode(image.data_ptr(), image.stride(0), w, h, int(quality), out.ctypes.data, C.byref(n),
                                          stream), "vr180_jpeg_encode")
    return out[: n.value].tobytes()
