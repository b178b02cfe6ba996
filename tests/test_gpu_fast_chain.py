"""GPU parity tests of the folded standard chain (csrc/chain_fast.cuh) that k_warp_tiled's prologue evaluates for
Normalize . EquirectangularEncoder . [Euclidean3DRotator] . [PolynomialScaler] . FisheyeDecoder("equidistant") .
Denormalize -- the chain apply_lr builds (remapper.py:23-59, :448-456).

The folded form differs from the reference's order of operations by a few float64 roundings (~1e-12 px); the float32
map and its 1/32-pixel quantisation must come out the same.  Checked two ways: against cv2.remap on the oracle's maps
(oracle/chain_np.py) at sizes the oracle finishes in seconds, and at full size (2 x 2048^2 .. 4096^2 coordinates per
eye) against the op-by-op evaluation of chain.cuh in the same kernel (`vr180_debug_set(1, 256)`), which the golden
tests pin to the reference.  Noise images: a coordinate that moves by 1/32 px changes the bilinear result.
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np

pytestmark = pytest.mark.gpu

Q1 = (0.9999093510664558, 0.00500054686470522, 0.01000109372941044, -0.00750082029705783)
Q2 = (0.97, -0.12, 0.2, 0.05)       # normalised by quaternion(): a 28 degree rotation
POLY = [0, 1, -0.02, 0.003]
OP_BY_OP = 256


@pytest.fixture()
def flags():
    lib = V._native.lib()
    yield lambda f: lib.vr180_debug_set(1, int(f))
    lib.vr180_debug_set(1, -1)


def _variants():
    enc, rot, poly, dec = V.EquirectangularEncoder, V.Euclidean3DRotator, V.PolynomialScaler, V.FisheyeDecoder
    return {
        "rot_poly": (lambda q: enc() * rot(V.quaternion(*q)) * poly(POLY) * dec("equidistant"),
                     lambda q: [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()), ("poly", POLY),
                                ("fisheye_dec", "equidistant")]),
        "base": (lambda q: enc() * dec("equidistant"), lambda q: [("equirect_enc", True), ("fisheye_dec", "equidistant")]),
        "poly_c0": (lambda q: enc() * poly([0.01, 0.9, 0.05]) * dec("equidistant"),   # theta = 0 keeps a radius: roll(0, 0) = 0
                    lambda q: [("equirect_enc", True), ("poly", [0.01, 0.9, 0.05]), ("fisheye_dec", "equidistant")]),
        "lat_x_rot": (lambda q: enc(is_latitude_y=False) * rot(V.quaternion(*q)) * dec("equidistant"),
                      lambda q: [("equirect_enc", False), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()),
                                 ("fisheye_dec", "equidistant")]),
        "neg_poly": (lambda q: enc() * rot(V.quaternion(*q)) * poly([0, -1, 0.1]) * dec("equidistant"),   # negative radii
                     lambda q: [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()),
                                ("poly", [0, -1, 0.1]), ("fisheye_dec", "equidistant")]),
    }


def _noise(seed, n, h, w):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8), rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("name", sorted(_variants()))
def test_fast_chain_matches_the_oracle(name, interp):
    """Every variant of the standard chain, 640 x 480 per eye from 500 x 520 sources (odd centre row: the optical axis
    falls on a pixel), per-eye maps (both chain slots), 5 frames (a multi-frame CTA), all four interpolations."""
    import torch

    make, ops = _variants()[name]
    hin, win, wout, hout = 500, 520, 640, 480
    ln, rn = _noise(3 + interp, 5, hin, win)
    wp = V.SbsWarper((make(Q1), make(Q2)), size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=255.0,
                     map_source="analytic")
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    ml = chain_np.get_map(ops(Q1), radius=255.0, size_input=(hin, win), size_output=(wout, hout))
    mr = chain_np.get_map(ops(Q2), radius=255.0, size_input=(hin, win), size_output=(wout, hout))
    for f in (0, 4):
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp), cv2.remap(rn[f], mr[0], mr[1], interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (name, interp, f, int((got[f] != want).sum()))


@pytest.mark.parametrize("name,n", [("rot_poly", 4096), ("base", 4096), ("lat_x_rot", 2048), ("poly_c0", 2048), ("neg_poly", 2048)])
def test_fast_chain_equals_the_op_by_op_chain(flags, name, n):
    """Full-size outputs (2 x n^2 coordinates, per-eye maps): the folded evaluation and the op-by-op evaluation give the
    same frame, byte for byte, on noise sources (bilinear: any coordinate that moved by 1/32 px shows)."""
    import torch

    make, _ = _variants()[name]
    hin = win = n // 2
    ln, rn = _noise(17, 2, hin, win)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    wp = V.SbsWarper((make(Q1), make(Q2)), size_input=(hin, win), size_output=(n, n), interpolation=1, radius=hin / 2,
                     map_source="analytic")
    flags(OP_BY_OP)
    ref = wp(left, right)
    flags(0)
    got = wp(left, right)
    assert int((got != ref).sum()) == 0
    assert int(ref.sum()) > 0


def test_fast_chain_per_frame_radius(flags):
    """radius="auto": the folded chain stops before Denormalize; frames with different disc radii."""
    import torch

    make, ops = _variants()["rot_poly"]
    hin = win = 384
    wout, hout = 640, 320
    rng = np.random.default_rng(5)
    frames = []
    for r in (180, 180, 150, 170, 170, 170):   # get_radius sees the disc
        img = rng.integers(16, 256, (hin, win, 3), dtype=np.uint8)
        yy, xx = np.ogrid[:hin, :win]
        img[(xx - win // 2) ** 2 + (yy - hin // 2) ** 2 > r * r] = 0
        frames.append(img)
    ln = np.stack(frames)
    rn = ln[::-1].copy()
    wp = V.SbsWarper(make(Q1), size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius="auto", map_source="analytic")
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    got = wp(left, right).cpu().numpy()
    flags(OP_BY_OP)
    assert np.array_equal(wp(left, right).cpu().numpy(), got)
    for f in range(len(frames)):
        radius = max(chain_np.get_radius(ln[f]), chain_np.get_radius(rn[f]))
        m = chain_np.get_map(ops(Q1), radius=radius, size_input=(hin, win), size_output=(wout, hout))
        want = np.concatenate([cv2.remap(ln[f], m[0], m[1], interpolation=1), cv2.remap(rn[f], m[0], m[1], interpolation=1)], axis=1)
        assert np.array_equal(got[f], want), (f, radius, int((got[f] != want).sum()))


def test_non_orthonormal_matrix_keeps_the_op_by_op_chain():
    """A `rot3` matrix that is not a rotation (only reachable by lowering to the C ABI by hand: here a 3 % shear-and-scale)
    makes (hypot(vx, vy), vz) leave the unit circle, so the folded theta would be wrong by percents: the host detects it
    (|R^T R - I| >= 1e-12) and the kernel evaluates arccos(vz) op by op, as the reference would with such a matrix."""
    import torch

    M = chain_np.quat_to_matrix(*Q2) @ np.array([[1.03, 0.01, 0.0], [0.0, 0.98, 0.0], [0.0, 0.02, 1.0]])

    @__import__("attrs").define()
    class Sheared(V.Euclidean3DRotator):
        def lower(self, shape=None, inverse=False):
            return [("rot3", M.reshape(-1).tolist())]

    t = V.EquirectangularEncoder() * Sheared(V.quaternion(*Q2)) * V.PolynomialScaler(POLY) * V.FisheyeDecoder("equidistant")
    ops = [("equirect_enc", True), ("rot3", M.ravel().tolist()), ("poly", POLY), ("fisheye_dec", "equidistant")]
    hin, win, wout, hout = 500, 520, 640, 480
    ln, rn = _noise(29, 3, hin, win)
    wp = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius=255.0, map_source="analytic")
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    m = chain_np.get_map(ops, radius=255.0, size_input=(hin, win), size_output=(wout, hout))
    assert np.isnan(m[0]).any()  # |vz| > 1 somewhere: np.arccos gives NaN there, and so must the kernel (border colour)
    for f in range(3):
        want = np.concatenate([cv2.remap(ln[f], m[0], m[1], interpolation=1), cv2.remap(rn[f], m[0], m[1], interpolation=1)], axis=1)
        assert np.array_equal(got[f], want), (f, int((got[f] != want).sum()))
