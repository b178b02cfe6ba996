"""GPU parity tests of the regime bench.py times: MANY frames per CTA of the tiled kernel (csrc/tiled.cu).

The launcher gives a CTA as many frames as the grid allows (`frames_per_cta`), which for the small outputs a test
can afford collapses to one frame per CTA -- the stage ring would never be re-filled, the mbarrier phases never flip
and the per-frame-radius loop never sees a second frame.  `vr180_debug_set(0, n)` (include/vr180_b200.h, test hook)
forces n frames per CTA, so that with 47 small frames every CTA runs 12-47 pipeline items through a ring of
S = 2-8 stages: the regime of the 64-frames-per-CTA benchmark launch.  Every frame of every case is compared with
cv2.remap (the reference's sampler, remapper.py:388-398) driven by the oracle's maps (oracle/chain_np.py), bit for bit.
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np

pytestmark = pytest.mark.gpu

HIN, WIN, WOUT, HOUT = 96, 128, 128, 64   # 4 x 2 bilinear tiles, 4 x 4 bicubic, 4 x 8 Lanczos4: every tile is full
N_FRAMES = 47                             # odd: two- and four-frame items end with phantom frames
QL = (0.9995, 0.012, -0.02, 0.015)
QR = (0.9995, -0.012, 0.02, -0.015)
POLY = [0, 1, 0.04]


@pytest.fixture()
def frames_per_cta():
    lib = V._native.lib()

    def set_(n):
        lib.vr180_debug_set(0, int(n))

    yield set_
    lib.vr180_debug_set(0, 0)


def _chain(q):
    return (V.EquirectangularEncoder() * V.Euclidean3DRotator(V.quaternion(*q)) * V.PolynomialScaler(POLY)
            * V.FisheyeDecoder("equidistant"))


def _ops(q):
    return [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()), ("poly", POLY),
            ("fisheye_dec", "equidistant")]


_MAPS: dict = {}


def _omap(q, radius):
    key = (q, radius)
    if key not in _MAPS:
        _MAPS[key] = chain_np.get_map(_ops(q), radius=radius, size_input=(HIN, WIN), size_output=(WOUT, HOUT))
    return _MAPS[key]


def _frames(seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (N_FRAMES, HIN, WIN, 3), dtype=np.uint8)  # no black surround: every tap carries signal


def _check(got, ln, rn, interp, per_eye, radii):
    for f in range(N_FRAMES):
        r = radii[f]
        if r != r:  # NaN radius: every coordinate NaN -> border colour
            assert not got[f].any(), ("nan frame", f)
            continue
        ml = _omap(QL, r)
        mr = _omap(QR, r) if per_eye else ml
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp),
                               cv2.remap(rn[f], mr[0], mr[1], interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, per_eye, f, r, int((got[f] != want).sum()))


@pytest.mark.parametrize("fpc", [24, 47])
@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("per_eye", [False, True])
def test_long_frame_loop_fixed_radius(frames_per_cta, interp, per_eye, fpc):
    """Fixed radius: FR = 2 (4 for bicubic) frames per item, ring refills + phase flips, phantom frames at the end of
    the odd chunk, shared map (2 views per CTA) and per-eye maps."""
    import torch

    frames_per_cta(fpc)
    radius = 70.0  # > HIN / 2: tiles straddle the top / bottom source edge (TMA zero fill)
    t = (_chain(QL), _chain(QR)) if per_eye else _chain(QL)
    ln, rn = _frames(1), _frames(2)
    out = torch.full((N_FRAMES + 1, HOUT, 2 * WOUT, 3), 99, dtype=torch.uint8, device="cuda")  # guard frame behind the batch
    V.SbsWarper(t, size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius=radius)(
        torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda(), out=out[:N_FRAMES])
    got = out.cpu().numpy()
    assert (got[N_FRAMES] == 99).all(), "a phantom frame was stored behind the batch"
    _check(got, ln, rn, interp, per_eye, [radius] * N_FRAMES)


@pytest.mark.parametrize("src", ["lut", "lut_fixed", "lut_packed"])
@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_long_frame_loop_lut_sources(frames_per_cta, interp, src):
    """The cached-LUT coordinate sources (float2 maps, fixed-point LUT, tile-packed LUT) through the same long frame
    loop."""
    import torch

    if src == "lut_fixed" and interp == 0:
        pytest.skip("the fixed-point LUT stores x * 32; INTER_NEAREST rounds x itself")
    frames_per_cta(47)
    radius = 61.5
    ln, rn = _frames(3), _frames(4)
    got = V.SbsWarper(_chain(QL), size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius=radius,
                      map_source=src)(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    _check(got, ln, rn, interp, False, [radius] * N_FRAMES)


def _radii(kind):
    if kind == "all_different":  # the per-pixel constants are rebuilt for every frame; every frame has its own rectangle
        base = [40.0, 70.25, -55.5, 63.0, 48.5, 90.0, -62.0, 57.75]
        return [base[i % len(base)] + 0.5 * (i // len(base)) for i in range(N_FRAMES)]
    if kind == "runs":  # a static rig with glitches: runs of equal radii, NaN (no transition found) in between
        pat = [60.5] * 5 + [float("nan")] * 3 + [60.5] * 2 + [44.0] * 7 + [float("nan")] + [71.5] * 6
        return [pat[i % len(pat)] for i in range(N_FRAMES)]
    if kind == "equal_chunk_then_mixed":  # chunk 0 (24 frames) takes the fixed-radius pipeline inside the DYN kernel
        return [58.5] * 24 + [58.5, 47.0] * 11 + [float("nan")]
    raise AssertionError(kind)


@pytest.mark.parametrize("kind", ["all_different", "runs", "equal_chunk_then_mixed"])
@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("per_eye", [False, True])
def test_long_frame_loop_per_frame_radius(frames_per_cta, interp, per_eye, kind):
    """Per-frame radius from device memory (DYN instantiation): > 1 frame per CTA, the stage origin hand-off
    (`s_org`), the "all radii equal" detection over a real chunk, runs of equal radii and NaN radii."""
    import torch

    frames_per_cta(24)
    radii = _radii(kind)
    t = (_chain(QL), _chain(QR)) if per_eye else _chain(QL)
    ln, rn = _frames(5), _frames(6)
    wp = V.SbsWarper(t, size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius="auto")
    got = wp(torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda(),
             radius=torch.tensor(radii, dtype=torch.float64, device="cuda")).cpu().numpy()
    _check(got, ln, rn, interp, per_eye, radii)


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("border", [(0, (7, 200, 90)), (1, 0), (2, 0), (3, 0), (4, 0), (0, 5)])
def test_tiled_kernel_with_every_border(frames_per_cta, interp, border):
    """Border modes / colours other than BORDER_CONSTANT(0) on the tiled kernel: tiles whose footprint stays inside the
    source are staged by TMA as usual, tiles that touch the source edge (a radius larger than the source makes many)
    take the per-pixel path with cv2's border rules; fixed radius and per-frame radius (incl. a NaN radius, which under
    BORDER_REPLICATE samples the corner pixel instead of a colour)."""
    import torch

    mode, value = border
    frames_per_cta(12)
    n = 12
    ln, rn = _frames(7)[:n], _frames(8)[:n]
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    t = _chain(QL)
    bv = value if isinstance(value, tuple) else (value, 0, 0)

    def want(f, r):
        if r != r:
            xm = np.full((HOUT, WOUT), np.nan, np.float32)
            ml = (xm, xm)
        else:
            ml = _omap(QL, r)
        return np.concatenate([cv2.remap(img[f], ml[0], ml[1], interpolation=interp, borderMode=mode, borderValue=bv)
                               for img in (ln, rn)], axis=1)

    radius = 75.0
    got = V.SbsWarper(t, size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius=radius,
                      boarder_mode=mode, boarder_value=value)(left, right).cpu().numpy()
    for f in range(n):
        assert np.array_equal(got[f], want(f, radius)), (interp, border, f)
    radii = [75.0, 40.0, float("nan"), 75.0, 75.0, 52.5, -66.0, float("nan"), 30.0, 75.0, 75.0, 75.0]
    got = V.SbsWarper(t, size_input=(HIN, WIN), size_output=(WOUT, HOUT), interpolation=interp, radius="auto",
                      boarder_mode=mode, boarder_value=value)(
        left, right, radius=torch.tensor(radii, dtype=torch.float64, device="cuda")).cpu().numpy()
    for f in range(n):
        assert np.array_equal(got[f], want(f, radii[f])), (interp, border, "dyn", f, radii[f])
