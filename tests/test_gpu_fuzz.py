"""Seeded random sweep of the device-resident warp (vr180_remap through SbsWarper) against the oracle.

Every case draws: a transformer chain (rotation, polynomial, zoom, fisheye mapping), per-eye or shared maps, input and
output sizes (odd widths, partial tiles, outputs smaller than one tile), interpolation, border mode and colour, batch
size and coordinate source.  The batch size and the source decide which kernel `vr180_remap` takes (csrc/tiled.cu:
batches; csrc/stream.cu: up to 12 (frame, eye) items per tile with a tile-packed LUT; csrc/kernels.cu: buffers the
tiled kernels cannot take), so the sweep crosses every dispatch boundary.  The expected frames are cv2.remap (the
reference's sampler, remapper.py:388-398) on the maps of oracle/chain_np.py (the reference's get_map restated),
concatenated like remapper.py:518 -- bit for bit.

One class of pixels is excluded: a projection's singular point falling exactly on output pixels (the antipode of the
stereographic mapping, 90 degrees off axis for the rectilinear one, e.g. the pole row; mostly reachable only with
output aspect ratios that run the latitude past the poles).  There one coordinate is ~1e17 px and the other the product of a ~1e-17 direction cosine and
that radius: NumPy's atan2 -> sin / cos round trip and the kernel's v / |v| form (csrc/chain.cuh) are both correctly
rounded evaluations of an expression with condition number 1e17 and differ by tens of pixels.  With the default constant
border such a pixel is the border colour either way; the reflecting / wrapping borders make the difference visible.
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np

pytestmark = pytest.mark.gpu

MAPPINGS = ["equidistant", "stereographic", "equisolid", "orthographic", "rectilinear"]
BORDERS = [cv2.BORDER_CONSTANT, cv2.BORDER_REPLICATE, cv2.BORDER_REFLECT, cv2.BORDER_WRAP, cv2.BORDER_REFLECT_101]


def _draw_chain(rng):
    """One eye's transformer and the oracle's op list for it."""
    t = V.EquirectangularEncoder()
    ops = [("equirect_enc", True)]
    if rng.random() < 0.7:
        rv = rng.normal(0, 0.03, 3)
        ang = float(np.linalg.norm(rv))
        q = (np.cos(ang / 2), *(np.sin(ang / 2) * rv / ang)) if ang > 0 else (1.0, 0.0, 0.0, 0.0)
        t = t * V.Euclidean3DRotator(V.quaternion(*q))
        ops.append(("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()))
    if rng.random() < 0.6:
        poly = [0.0, 1.0, float(rng.normal(0, 0.03)), float(rng.normal(0, 0.005))]
        t = t * V.PolynomialScaler(poly)
        ops.append(("poly", poly))
    if rng.random() < 0.3:
        z = float(rng.uniform(0.8, 1.3))
        t = t * V.ZoomTransformer(z)
        ops.append(("zoom", z))
    mapping = MAPPINGS[int(rng.integers(0, len(MAPPINGS)))] if rng.random() < 0.4 else "equidistant"
    t = t * V.FisheyeDecoder(mapping)
    ops.append(("fisheye_dec", mapping))
    return t, ops


@pytest.mark.parametrize("seed", range(96))
def test_random_case_matches_oracle(seed):
    import torch

    rng = np.random.default_rng(1000 + seed)
    hin, win = int(rng.integers(40, 260)), int(rng.integers(40, 300))
    wout = int(rng.choice([16, 32, 64, 96, 100, 128, 136, 160, 208, 256]))
    hout = int(rng.choice([8, 16, 24, 32, 40, 64, 72, 96, 104, 128]))
    interp = int(rng.choice([0, 1, 1, 2, 4]))
    border = BORDERS[int(rng.integers(0, len(BORDERS)))] if rng.random() < 0.5 else cv2.BORDER_CONSTANT
    value = tuple(int(v) for v in rng.integers(0, 256, 3)) if rng.random() < 0.3 else (0, 0, 0)
    n_frames = int(rng.choice([1, 1, 2, 3, 5, 7, 14, 23]))
    per_eye = bool(rng.random() < 0.5)
    source = str(rng.choice(["auto", "analytic", "lut", "lut_fixed", "lut_packed", "lut_packed"]))
    if source == "lut_fixed" and interp == 0:
        source = "lut"
    radius = float(rng.uniform(0.35, 0.8) * min(hin, win))
    tl, ops_l = _draw_chain(rng)
    tr, ops_r = _draw_chain(rng) if per_eye else (tl, ops_l)
    ln = rng.integers(0, 256, (n_frames, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n_frames, hin, win, 3), dtype=np.uint8)
    wp = V.SbsWarper((tl, tr) if per_eye else tl, size_input=(hin, win), size_output=(wout, hout), interpolation=interp,
                     radius=radius, map_source=source, boarder_mode=border, boarder_value=value)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    got = wp(left, right).cpu().numpy()
    got2 = wp(left, right).cpu().numpy()  # "auto" switches to its cached LUT on the second small call
    ml = chain_np.get_map(ops_l, radius=radius, size_input=(hin, win), size_output=(wout, hout))
    mr = chain_np.get_map(ops_r, radius=radius, size_input=(hin, win), size_output=(wout, hout)) if per_eye else ml
    case = dict(seed=seed, size_in=(hin, win), size_out=(wout, hout), interp=interp, border=border, value=value,
                n_frames=n_frames, per_eye=per_eye, source=source)
    singular = np.concatenate([(np.abs(m[0]) > 1e9) | (np.abs(m[1]) > 1e9) for m in (ml, mr)], axis=1)
    assert singular.mean() <= 0.15, (case, int(singular.sum()))  # a pole row or two of the rectilinear mapping at most
    for f in range(n_frames):
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp, borderMode=border, borderValue=value),
                               cv2.remap(rn[f], mr[0], mr[1], interpolation=interp, borderMode=border, borderValue=value)],
                              axis=1)
        want[singular] = got[f][singular]
        assert np.array_equal(got[f], want), (case, f, int((got[f] != want).sum()))
        assert np.array_equal(got2[f], want), (case, f, "second call", int((got2[f] != want).sum()))
