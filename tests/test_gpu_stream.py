"""GPU parity tests of the tile-streaming kernel (csrc/stream.cu): launches of one or two frames with tile-packed LUTs.

A persistent CTA walks many tiles; its producer warp fetches the source rectangles of the next tiles (bounding box
from the packed tile header) while the sampling warps are still on the current one.  `vr180_debug_set(3, n)` shrinks
the grid to n CTAs so that a small test output pushes dozens to hundreds of tiles through ONE CTA: the slot ring and
the out-buffer ring wrap many times, items of consecutive tiles have different rectangles, and non-staged tiles
(unpackable, outside the source under a non-zero border) are interleaved with staged ones.  Every frame is compared
with cv2.remap (the reference's sampler, remapper.py:388-398) on the oracle's maps (oracle/chain_np.py), bit for bit,
and with the batch kernel (csrc/tiled.cu) forced onto the same request.
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np

pytestmark = pytest.mark.gpu

HIN, WIN, WOUT, HOUT = 200, 240, 256, 160   # 8 x 5 bilinear tiles, 8 x 10 bicubic, 8 x 20 Lanczos4
QL = (0.9995, 0.012, -0.02, 0.015)
QR = (0.9995, -0.012, 0.02, -0.015)
POLY = [0, 1, 0.04]


@pytest.fixture()
def hooks():
    lib = V._native.lib()

    def set_(grid=0, flags=-1):
        lib.vr180_debug_set(3, int(grid))
        lib.vr180_debug_set(1, int(flags))

    yield set_
    lib.vr180_debug_set(3, 0)
    lib.vr180_debug_set(1, -1)


def _chain(q):
    return (V.EquirectangularEncoder() * V.Euclidean3DRotator(V.quaternion(*q)) * V.PolynomialScaler(POLY)
            * V.FisheyeDecoder("equidistant"))


def _ops(q):
    return [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()), ("poly", POLY),
            ("fisheye_dec", "equidistant")]


def _want(ln, rn, interp, per_eye, radius, border=cv2.BORDER_CONSTANT, value=(0, 0, 0)):
    ml = chain_np.get_map(_ops(QL), radius=radius, size_input=(HIN, WIN), size_output=(WOUT, HOUT))
    mr = chain_np.get_map(_ops(QR), radius=radius, size_input=(HIN, WIN), size_output=(WOUT, HOUT)) if per_eye else ml
    return np.stack([np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp, borderMode=border, borderValue=value),
                                     cv2.remap(rn[f], mr[0], mr[1], interpolation=interp, borderMode=border, borderValue=value)],
                                    axis=1) for f in range(len(ln))])


def _launches():
    return int(V._native.lib().vr180_launch_count())


@pytest.mark.parametrize("grid", [0, 3])
@pytest.mark.parametrize("n_frames", [1, 2])
@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("per_eye", [False, True])
def test_stream_matches_oracle_and_batch_kernel(hooks, interp, per_eye, n_frames, grid):
    """One and two stereo pairs, shared map (both eyes per tile) and per-eye maps (two map groups), the default zero
    border (tiles straddling the source edge stay staged: radius > HIN / 2), default grid and 3 CTAs for 40-320 tiles."""
    import torch

    rng = np.random.default_rng(7 * interp + n_frames)
    ln = rng.integers(0, 256, (n_frames, HIN, WIN, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n_frames, HIN, WIN, 3), dtype=np.uint8)
    t = (_chain(QL), _chain(QR)) if per_eye else _chain(QL)
    radius = 130.0
    wp = V.SbsWarper(t, size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius=radius,
                     map_source="lut_packed")
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    hooks(grid, 2)  # streaming kernel whenever eligible
    out = torch.full((n_frames + 1, HOUT, 2 * WOUT, 3), 99, dtype=torch.uint8, device="cuda")
    wp(left, right, out=out[:n_frames])
    got = out.cpu().numpy()
    assert (got[n_frames] == 99).all()
    want = _want(ln, rn, interp, per_eye, radius)
    assert np.array_equal(got[:n_frames], want), (interp, per_eye, n_frames, int((got[:n_frames] != want).sum()))
    hooks(0, 4)  # never: the batch kernel on the same request
    assert np.array_equal(wp(left, right).cpu().numpy(), want)


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("border", ["colour", cv2.BORDER_REPLICATE, cv2.BORDER_REFLECT_101, cv2.BORDER_WRAP])
def test_stream_borders(hooks, interp, border):
    """Any border other than constant zero: tiles whose footprint leaves the source take the per-pixel path between
    staged tiles of the same CTA (2 CTAs for the whole output)."""
    import torch

    rng = np.random.default_rng(11 + interp)
    ln = rng.integers(0, 256, (1, HIN, WIN, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (1, HIN, WIN, 3), dtype=np.uint8)
    mode, value = (cv2.BORDER_CONSTANT, (17, 200, 3)) if border == "colour" else (border, (0, 0, 0))
    wp = V.SbsWarper(_chain(QL), size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius=130.0,
                     map_source="lut_packed", boarder_mode=mode, boarder_value=value)
    hooks(2, 2)
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    want = _want(ln, rn, interp, False, 130.0, mode, value)
    assert np.array_equal(got, want), (interp, border, int((got != want).sum()))


def test_stream_is_the_default_for_a_single_pair(hooks):
    """Without hooks a single pair with a packed LUT takes the streaming kernel (one launch, same result as the batch
    kernel); a 5-pair batch of the same plan gives the same frames (streamed too: 10 items per tile)."""
    import torch

    rng = np.random.default_rng(5)
    ln = rng.integers(0, 256, (5, HIN, WIN, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (5, HIN, WIN, 3), dtype=np.uint8)
    wp = V.SbsWarper(_chain(QL), size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=1, radius=130.0,
                     map_source="lut_packed")
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    want = _want(ln, rn, 1, False, 130.0)
    wp(left[:1], right[:1])  # builds the maps and the LUT
    n0 = _launches()
    got1 = wp(left[:1], right[:1]).cpu().numpy()
    assert _launches() - n0 == 1
    assert np.array_equal(got1, want[:1])
    assert np.array_equal(wp(left, right).cpu().numpy(), want)


def test_stream_long_run_through_one_cta(hooks):
    """640 bilinear tiles (1024 x 640 output) x 2 eyes through ONE CTA: 1280 items, the 3-slot ring wraps 426 times."""
    import torch

    hin, win, wout, hout = 300, 360, 1024, 640
    rng = np.random.default_rng(9)
    ln = rng.integers(0, 256, (1, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (1, hin, win, 3), dtype=np.uint8)
    wp = V.SbsWarper(_chain(QL), size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius=170.0,
                     map_source="lut_packed")
    hooks(1, 2)
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    m = chain_np.get_map(_ops(QL), radius=170.0, size_input=(hin, win), size_output=(wout, hout))
    want = np.concatenate([cv2.remap(ln[0], m[0], m[1], interpolation=1), cv2.remap(rn[0], m[0], m[1], interpolation=1)], axis=1)
    assert np.array_equal(got[0], want), int((got[0] != want).sum())


def test_stream_more_unstaged_tiles_than_the_list_holds(hooks):
    """BORDER_REPLICATE with a radius far beyond the source: most of the 640 tiles leave the source and are left to the
    per-pixel gather.  One CTA notes at most 128 of them before it stops streaming, gathers them and resumes."""
    import torch

    hin, win, wout, hout = 300, 360, 1024, 640
    rng = np.random.default_rng(10)
    ln = rng.integers(0, 256, (1, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (1, hin, win, 3), dtype=np.uint8)
    wp = V.SbsWarper(_chain(QL), size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius=420.0,
                     map_source="lut_packed", boarder_mode=cv2.BORDER_REPLICATE)
    hooks(1, 2)
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    m = chain_np.get_map(_ops(QL), radius=420.0, size_input=(hin, win), size_output=(wout, hout))
    ix, iy = np.floor(m[0]), np.floor(m[1])
    outside = ((ix < 0) | (ix + 1 > win - 1) | (iy < 0) | (iy + 1 > hin - 1)).reshape(hout // 32, 32, wout // 32, 32).any(axis=(1, 3))
    assert 128 < outside.sum() < outside.size, outside.sum()  # the list overflows, and staged tiles remain
    want = np.concatenate([cv2.remap(ln[0], m[0], m[1], interpolation=1, borderMode=cv2.BORDER_REPLICATE),
                           cv2.remap(rn[0], m[0], m[1], interpolation=1, borderMode=cv2.BORDER_REPLICATE)], axis=1)
    assert np.array_equal(got[0], want), int((got[0] != want).sum())


def test_auto_map_source_switches_to_the_cached_lut_on_the_second_small_call(hooks):
    """SbsWarper's default map_source="auto": batches and the first small call are analytic; from the second small call
    on the plan serves one- / two-pair calls from its tile-packed LUT (streaming kernel); auto radius stays analytic."""
    import torch

    rng = np.random.default_rng(12)
    ln = rng.integers(0, 256, (11, HIN, WIN, 3), dtype=np.uint8)  # 11 pairs = 22 (frame, eye) items per tile: a batch
    rn = rng.integers(0, 256, (11, HIN, WIN, 3), dtype=np.uint8)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    want = _want(ln, rn, 1, False, 130.0)
    wp = V.SbsWarper(_chain(QL), size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=1, radius=130.0)
    assert wp.map_source == "auto"
    assert np.array_equal(wp(left[:1], right[:1]).cpu().numpy(), want[:1]) and wp._packed is None  # analytic
    assert np.array_equal(wp(left, right).cpu().numpy(), want) and wp._packed is None              # a batch: analytic
    assert np.array_equal(wp(left[1:3], right[1:3]).cpu().numpy(), want[1:3]) and wp._packed is not None
    n0 = _launches()
    assert np.array_equal(wp(left[3:4], right[3:4]).cpu().numpy(), want[3:4])
    assert _launches() - n0 == 1
    auto = V.SbsWarper(_chain(QL), size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=1, radius="auto")
    for _ in range(3):
        auto(left[:1], right[:1])
    assert auto._packed is None
