"""Seeded random sweep of the host-array API (apply / lr_frame / lr_frames -> vr180_ctx_run) against the oracle.

The host pipeline (csrc/pipeline.cu) re-lays every job out for the tiled kernels: row pitches rounded up to 16 bytes,
the right eye at a 16-byte aligned column (downloaded as column segments when W*3 is not a multiple of 16), pageable
arrays packed into a pinned ring, strided views (the reference passes views of an SBS image, remapper.py:455-456),
scattered frame lists, chunks of a few frames.  Each case draws sizes (odd widths included), a chain, interpolation,
border, radius mode ("auto" / "max" / a number, with get_radius on the device), shared or per-eye transformers, 1-9
frames, views or contiguous arrays, and compares every output with the reference semantics restated by the oracle:
get_radius_smart (remapper.py:62-90) -> get_map -> cv2.remap per eye -> np.concatenate (remapper.py:388-398, :518).
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np
from tests.conftest import disc_frame
from tests.test_gpu_fuzz import BORDERS, _draw_chain

pytestmark = pytest.mark.gpu


def _frame(rng, h, w, as_view):
    """A fisheye-like frame (disc with a black surround so that get_radius finds its transitions); optionally a strided
    view into a wider array, as the reference's split of an SBS image produces."""
    img = disc_frame(h, w, seed=int(rng.integers(0, 1 << 30)), margin=int(rng.integers(2, 9)))
    if not as_view:
        return img
    wide = np.zeros((h, w + int(rng.integers(1, 40)), 3), np.uint8)
    x0 = int(rng.integers(0, wide.shape[1] - w + 1))
    wide[:, x0:x0 + w] = img
    return wide[:, x0:x0 + w]


def _radius_of(mode, images):
    if mode == "auto":
        return max(chain_np.get_radius(im) for im in images)
    if mode == "max":
        return min(images[0].shape[0] / 2, images[0].shape[1] / 2)
    return float(mode)


@pytest.mark.parametrize("seed", range(40))
def test_random_host_case_matches_oracle(seed):
    rng = np.random.default_rng(5000 + seed)
    hin, win = int(rng.integers(48, 200)), int(rng.integers(48, 240))
    wout = int(rng.choice([32, 50, 64, 77, 100, 128, 160]))
    hout = int(rng.choice([16, 32, 40, 64, 96]))
    interp = int(rng.choice([0, 1, 1, 2, 4]))
    border = BORDERS[int(rng.integers(0, len(BORDERS)))] if rng.random() < 0.4 else cv2.BORDER_CONSTANT
    value = int(rng.integers(0, 256)) if rng.random() < 0.3 else 0
    n = int(rng.choice([1, 1, 2, 3, 5, 9]))
    per_eye = bool(rng.random() < 0.4)
    mode = rng.choice(["auto", "max", "number"])
    mode = float(rng.uniform(0.35, 0.7) * min(hin, win)) if mode == "number" else str(mode)
    as_view = bool(rng.random() < 0.4)
    tl, ops_l = _draw_chain(rng)
    tr, ops_r = _draw_chain(rng) if per_eye else (tl, ops_l)
    lefts = [_frame(rng, hin, win, as_view) for _ in range(n)]
    rights = [_frame(rng, hin, win, as_view) for _ in range(n)]
    t = (tl, tr) if per_eye else tl
    kw = dict(size_output=(wout, hout), interpolation=interp, boarder_mode=border, boarder_value=value, radius=mode)
    case = dict(seed=seed, size_in=(hin, win), size_out=(wout, hout), interp=interp, border=border, value=value, n=n,
                per_eye=per_eye, radius=mode, view=as_view)

    def remap(img, ops, radius):
        m = chain_np.get_map(ops, radius=radius, size_input=(hin, win), size_output=(wout, hout))
        out = cv2.remap(img, m[0], m[1], interpolation=interp, borderMode=border, borderValue=value)
        return out, (np.abs(m[0]) > 1e9) | (np.abs(m[1]) > 1e9)

    # lr_frames: every pair exactly as apply_lr would produce it (per-eye tuple: a radius per eye, remapper.py:460-473)
    got = V.lr_frames(t, lefts, rights, **kw)
    assert len(got) == n
    for i in range(n):
        if per_eye:
            (wl, sl), (wr, sr) = remap(lefts[i], ops_l, _radius_of(mode, [lefts[i]])), remap(rights[i], ops_r, _radius_of(mode, [rights[i]]))
        else:
            r = _radius_of(mode, [lefts[i], rights[i]])
            (wl, sl), (wr, sr) = remap(lefts[i], ops_l, r), remap(rights[i], ops_r, r)
        want, sing = np.concatenate([wl, wr], axis=1), np.concatenate([sl, sr], axis=1)
        want[sing] = got[i][sing]
        assert got[i].shape == want.shape and np.array_equal(got[i], want), (case, "lr_frames", i, int((got[i] != want).sum()))

    # apply: ONE radius (max over all images for "auto", remapper.py:82-84) and ONE map for the list
    images = lefts + rights
    out = V.apply(tl, in_paths=images, **kw)
    r = _radius_of(mode, images)
    assert len(out) == 2 * n
    for i, img in enumerate(images):
        want, sing = remap(img, ops_l, r)
        want[sing] = out[i][sing]
        assert np.array_equal(out[i], want), (case, "apply", i, int((out[i] != want).sum()))
