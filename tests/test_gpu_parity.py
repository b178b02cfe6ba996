"""GPU parity tests (-m gpu, run on the B200 box): the CUDA path through the C ABI against the golden vectors of
the unmodified reference, against the oracle on seeded inputs, and through size-independent properties at the
full BASELINE.json sizes.  Nothing here reads /root/reference.

Tolerances (BASELINE.json north_star): maps within 1e-4 px of the reference chain (asserted as: bit-identical
float32 except for at most a vanishing number of 1-ulp round-to-nearest flips); sampled uint8 pixels BIT-EXACT
for every interpolation whenever the maps are identical.
"""
from __future__ import annotations

import cv2
import numpy as np
import pytest

import vr180_convert_b200 as V
from oracle import chain_np, remap_np
from tests.conftest import disc_frame

pytestmark = pytest.mark.gpu

NS = {k: getattr(V, k) for k in V.__all__}
NS["np"] = np


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(2**31) - ai, ai)
    bi = np.where(bi < 0, -(2**31) - bi, bi)
    return np.abs(ai - bi)


def _assert_maps_close(got, want, what, max_flip_frac=1e-5, singular_ok=0):
    """`singular_ok`: number of elements allowed to differ arbitrarily -- only used for chains with a coordinate
    singularity ON the pixel grid (EquirectangularDecoder at the poles: longitude = atan2(~1e-16, ~1e-17) is pure
    rounding noise in the reference; any longitude is the same point of the sphere)."""
    assert got.shape == want.shape and got.dtype == np.float32, what
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    if singular_ok:
        bad = ~np.isclose(got, want, rtol=0, atol=1e-3, equal_nan=True)
        assert bad.sum() <= singular_ok, (what, int(bad.sum()))
        got = np.where(bad, want, got)
    # a NaN appears where float64 rounding pushed |v_z| just above 1; allow the GPU's
    # libm to disagree on a vanishing number of such singular pixels
    assert (nan_g != nan_w).sum() <= max(1, int(want.size * 1e-6)), what
    ok = ~(nan_g | nan_w)
    d = np.abs(got[ok].astype(np.float64) - want[ok].astype(np.float64))
    tol = 1e-4 + np.spacing(np.abs(want[ok]).astype(np.float32)).astype(np.float64)
    assert (d <= tol).all(), (what, float(d.max()))
    # "same float32": differences are at most one float32 ulp (measured at magnitude >= 1 px; coordinates that
    # are ~0 in the reference only by cancellation noise have no meaningful ulp) and vanishingly rare
    one_ulp = np.spacing(np.maximum(np.abs(want[ok]), 1.0).astype(np.float32)).astype(np.float64)
    assert (d <= one_ulp).all(), (what, float((d / one_ulp).max()))
    flips = d > 0.25 * one_ulp
    assert flips.mean() <= max_flip_frac + 1.0 / max(flips.size, 1), (what, float(flips.mean()))


# ---------------------------------------------------------------------------------------------------------
# (1) analytic maps  vs  reference get_map
# ---------------------------------------------------------------------------------------------------------
def test_get_map_matches_reference_golden(golden_maps):
    z, meta = golden_maps
    for name, case in meta["cases"].items():
        t = eval(case["expr"], NS)  # noqa: S307
        for si, (size_out, size_in, radius) in enumerate(meta["shapes"]):
            xm, ym = V.get_map(t, radius=radius, size_input=tuple(size_in), size_output=tuple(size_out))
            # small maps: tolerate one flipped element per map
            sing = 2 if any(o[0] == "equirect_dec" for o in case["ops"]) else 0
            _assert_maps_close(xm, z[f"{name}/{si}/x"], (name, si, "x"), max_flip_frac=2e-3, singular_ok=sing)
            _assert_maps_close(ym, z[f"{name}/{si}/y"], (name, si, "y"), max_flip_frac=2e-3, singular_ok=sing)


def test_get_map_survey_known_answers():
    xm, ym = V.get_map(V.EquirectangularEncoder() * V.FisheyeDecoder("equidistant"), radius=128.0,
                       size_input=(256, 256), size_output=(256, 256))
    assert (xm[0, 0], ym[0, 0]) == (128.0, 0.0) and (xm[128, 128], ym[128, 128]) == (128.0, 128.0)
    assert xm[37, 201].view(np.uint32) == 0x432585C3 and ym[37, 201].view(np.uint32) == 0x41EC3D23
    assert xm[255, 255].view(np.uint32) == 0x4301920C and ym[255, 255].view(np.uint32) == 0x437FFA64


@pytest.mark.parametrize("name,n", [("base", 2048), ("rot_poly", 4096)])
def test_get_map_fullsize_samples(golden_fullsize, golden_maps, name, n):
    """cfg2 / cfg3 sizes: sparse samples + one full row/column of the reference's maps."""
    _, meta = golden_maps
    t = eval(meta["cases"][name]["expr"], NS)  # noqa: S307
    xm, ym = V.get_map(t, radius=n / 2, size_input=(n, n), size_output=(n, n))
    z = golden_fullsize
    idx = z[f"{name}/{n}/idx"]
    _assert_maps_close(xm[idx[:, 0], idx[:, 1]], z[f"{name}/{n}/x"], (name, "x samples"), 1e-3)
    _assert_maps_close(ym[idx[:, 0], idx[:, 1]], z[f"{name}/{n}/y"], (name, "y samples"), 1e-3)
    _assert_maps_close(np.stack([xm[n // 3], ym[n // 3]]), z[f"{name}/{n}/row"], (name, "row"), 1e-3)
    _assert_maps_close(np.stack([xm[:, n // 5], ym[:, n // 5]]), z[f"{name}/{n}/col"], (name, "col"), 1e-3)


def test_get_map_vs_oracle_whole_map():
    """Whole 1024^2 map against the oracle chain on this host (rotated + polynomial chain)."""
    q = (0.9999093510664558, 0.00500054686470522, 0.01000109372941044, -0.00750082029705783)
    t = (V.EquirectangularEncoder() * V.Euclidean3DRotator(V.quaternion(*q)) * V.PolynomialScaler([0, 1, -0.02, 0.003])
         * V.FisheyeDecoder("equidistant"))
    ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()),
           ("poly", [0, 1, -0.02, 0.003]), ("fisheye_dec", "equidistant")]
    n = 1024
    xm, ym = V.get_map(t, radius=n / 2, size_input=(n, n), size_output=(n, n))
    wx, wy = chain_np.get_map(ops, radius=n / 2, size_input=(n, n), size_output=(n, n))
    _assert_maps_close(xm, wx, "x")
    _assert_maps_close(ym, wy, "y")


# ---------------------------------------------------------------------------------------------------------
# (2) sampling  vs  cv2.remap
# ---------------------------------------------------------------------------------------------------------
def test_remap_matches_cv2_golden(golden_remap):
    g = golden_remap
    src, xm, ym = g["src"], g["xmap"], g["ymap"]
    for interp in (0, 1, 2, 4):
        for bm in (0, 1, 2, 3, 4):
            for bi, bv in enumerate((0, 7, (3, 200, 90))):
                got = V.remap_maps(src, xm, ym, interpolation=interp, border_mode=bm, border_value=bv)
                assert np.array_equal(got, g[f"out/{interp}/{bm}/{bi}"]), (interp, bm, bi)


@pytest.mark.parametrize("channels", [1, 3, 4])
@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_remap_random_maps_vs_oracle(channels, interp):
    rng = np.random.default_rng(100 * channels + interp)
    src = rng.integers(0, 256, (61, 83, channels), dtype=np.uint8)
    if channels == 1:
        src = src[:, :, 0]
    h, w = 77, 130
    xm = (rng.random((h, w)) * 100 - 9).astype(np.float32)
    ym = (rng.random((h, w)) * 80 - 9).astype(np.float32)
    xm[0, :6] = [np.nan, np.inf, -np.inf, 1e9, -7e7, 6e7]
    ym[1, :6] = [np.nan, np.inf, -np.inf, 1e9, -7e7, 6e7]
    xm[2] = np.floor(xm[2]) + 1 / 64  # exact .5 ties of x*32
    for bm in (0, 1, 2, 3, 4):
        bv = (5, 60, 200, 9)[:channels] if channels > 1 else 5
        want = remap_np.remap(src, xm, ym, interp, bm, bv)
        got = V.remap_maps(src, xm, ym, interpolation=interp, border_mode=bm, border_value=bv)
        assert np.array_equal(got, want), (channels, interp, bm)
        assert np.array_equal(want, cv2.remap(src, xm, ym, interpolation=interp, borderMode=bm, borderValue=bv))


def test_remap_strided_view_and_empty_edges():
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, (50, 120, 3), dtype=np.uint8)
    view = big[:, 60:]  # row-strided half, as remapper.py:455-456 produces
    xm = (rng.random((40, 33)) * 70 - 5).astype(np.float32)
    ym = (rng.random((40, 33)) * 60 - 5).astype(np.float32)
    assert np.array_equal(V.remap_maps(view, xm, ym, interpolation=1), cv2.remap(view, xm, ym, interpolation=1))
    # 1-pixel source and 1x1 destination
    one = np.full((1, 1, 3), 77, np.uint8)
    m0 = np.zeros((1, 1), np.float32)
    for interp in (0, 1, 2, 4):
        assert np.array_equal(V.remap_maps(one, m0, m0, interpolation=interp, border_mode=1),
                              cv2.remap(one, m0, m0, interpolation=interp, borderMode=1))


# ---------------------------------------------------------------------------------------------------------
# (3) get_radius
# ---------------------------------------------------------------------------------------------------------
def test_get_radius_golden(golden_radius):
    for key, want in golden_radius.items():
        if key[0].isdigit():
            h, w, r = map(int, key.split("x"))
            yy, xx = np.mgrid[:h, :w]
            img = np.where(((xx - w // 2) ** 2 + (yy - h // 2) ** 2 <= r * r)[..., None], 200, 0).astype(np.uint8)
            img = np.repeat(img, 3, axis=2)
            assert V.get_radius(img) == want, key
    for img in (np.zeros((64, 64, 3), np.uint8), np.full((64, 64, 3), 255, np.uint8)):
        with pytest.raises(IndexError):
            V.get_radius(img)


def test_get_radius_random_lines_vs_oracle():
    rng = np.random.default_rng(11)
    for trial in range(20):
        h, w = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        img = rng.integers(0, 40, (h, w, 3), dtype=np.uint8)  # values straddle the threshold sum of 30
        thr = int(rng.integers(5, 15))
        try:
            want = chain_np.get_radius(img, threshold=thr)
        except IndexError:
            with pytest.raises(IndexError):
                V.get_radius(img, threshold=thr)
            continue
        assert V.get_radius(img, threshold=thr) == want, (trial, h, w)


# ---------------------------------------------------------------------------------------------------------
# (4) apply / apply_lr end to end  vs  the reference's outputs
# ---------------------------------------------------------------------------------------------------------
def test_apply_lr_matches_reference_golden(golden_apply, golden_maps, tmp_path):
    g = golden_apply
    _, meta = golden_maps
    card = g["card"]
    left, right = card[:, :128], card[:, 128:]
    for key in g.files:
        if not key.startswith("lr/"):
            continue
        _, name, interp, radius = key.split("/")
        radius = radius if radius in ("max", "auto") else float(radius)
        t = eval(meta["cases"][name]["expr"], NS)  # noqa: S307
        out = tmp_path / "o.png"
        V.apply_lr(t, left_path=left, right_path=right, out_path=out, size_output=(96, 80),
                   interpolation=int(interp), radius=radius)
        got = cv2.imread(str(out))
        assert got.shape == g[key].shape == (80, 192, 3)
        assert np.array_equal(got, g[key]), key
    tl = eval(meta["cases"]["rot_poly"]["expr"], NS)  # noqa: S307
    tr = eval(meta["cases"]["rot_nonunit"]["expr"], NS)  # noqa: S307
    got = V.lr_frame((tl, tr), left, right, size_output=(96, 80), interpolation=1, radius="auto")
    assert np.array_equal(got, g["lr_tuple/rot_poly+rot_nonunit/1/auto"])
    imgs = V.apply(eval(meta["cases"]["fe_poly"]["expr"], NS), in_paths=[card, card[::-1].copy()],  # noqa: S307
                   size_output=(72, 72), interpolation=1, radius="max", boarder_value=9)
    assert np.array_equal(imgs[0], g["apply/fe_poly/0"]) and np.array_equal(imgs[1], g["apply/fe_poly/1"])


def test_apply_lr_same_path_splits_halves(tmp_path, golden_apply):
    """remapper.py:448-456: identical paths -> one file split into halves."""
    card = golden_apply["card"]
    src = tmp_path / "sbs.png"
    cv2.imwrite(str(src), card)
    out = tmp_path / "out.png"
    t = V.EquirectangularEncoder() * V.FisheyeDecoder("equidistant")
    V.apply_lr(t, left_path=src, right_path=src, out_path=out, size_output=(64, 64), interpolation=1, radius="max")
    want = V.lr_frame(t, card[:, :128], card[:, 128:], size_output=(64, 64), interpolation=1, radius="max")
    assert np.array_equal(cv2.imread(str(out)), want)


def test_user_defined_transformer_goes_through_lut_path():
    class Mine(V.PolarRollTransformer):  # README.md:204-219
        def transform_polar(self, theta, roll, **kwargs):
            return theta**0.98 + theta**1.01, roll

    t = V.EquirectangularEncoder() * Mine() * V.FisheyeDecoder("equidistant")
    img = disc_frame(96, 96, seed=4)
    got = V.apply(t, in_paths=img, size_output=(64, 64), interpolation=1, radius=40.0)[0]
    xm, ym = np.meshgrid(np.arange(64), np.arange(64))
    fx, fy = (V.NormalizeTransformer() * t * V.DenormalizeTransformer(scale=(40.0, 40.0), center=(48, 48))).transform(xm, ym)
    want = cv2.remap(img, fx.astype(np.float32), fy.astype(np.float32), interpolation=1)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------------
# (5) batched device-resident path (SbsWarper): analytic == LUT == fixed LUT == oracle
# ---------------------------------------------------------------------------------------------------------
def _frames(n, h, w, seed0=0):
    return np.stack([disc_frame(h, w, seed=seed0 + i) for i in range(n)])


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_sbs_warper_sources_agree_and_match_oracle(interp):
    import torch

    n, hin, win, wout, hout = 5, 160, 160, 128, 96
    q = (0.9999093510664558, 0.00500054686470522, 0.01000109372941044, -0.00750082029705783)
    t = (V.EquirectangularEncoder() * V.Euclidean3DRotator(V.quaternion(*q)) * V.PolynomialScaler([0, 1, -0.02, 0.003])
         * V.FisheyeDecoder("equidistant"))
    left = torch.from_numpy(_frames(n, hin, win, 0)).cuda()
    right = torch.from_numpy(_frames(n, hin, win, 100)).cuda()
    outs = {}
    for src in ("analytic", "lut", "lut_fixed", "lut_packed"):
        if src == "lut_fixed" and interp == 0:
            continue
        wp = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=80.0,
                         map_source=src)
        outs[src] = wp(left, right).cpu().numpy()
    for src in outs:
        assert np.array_equal(outs[src], outs["analytic"]), src
    # oracle: reference chain restatement + cv2.remap per eye + concatenate
    ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q).ravel().tolist()),
           ("poly", [0, 1, -0.02, 0.003]), ("fisheye_dec", "equidistant")]
    xm, ym = chain_np.get_map(ops, radius=80.0, size_input=(hin, win), size_output=(wout, hout))
    ln, rn = left.cpu().numpy(), right.cpu().numpy()
    for f in range(n):
        want = np.concatenate([cv2.remap(ln[f], xm, ym, interpolation=interp),
                               cv2.remap(rn[f], xm, ym, interpolation=interp)], axis=1)
        assert np.array_equal(outs["analytic"][f], want), (interp, f)


def test_sbs_warper_per_eye_tuple_and_auto_radius():
    import torch

    n, hin, win, wout, hout = 3, 200, 200, 96, 96
    tl = V.EquirectangularEncoder() * V.Euclidean3DRotator(V.from_rotation_vector([0.01, 0.02, -0.015])) * V.FisheyeDecoder("equidistant")
    tr = V.EquirectangularEncoder() * V.Euclidean3DRotator(V.from_rotation_vector([-0.01, -0.02, 0.015])) * V.FisheyeDecoder("equidistant")
    frames_l, frames_r = _frames(n, hin, win, 0), _frames(n, hin, win, 50)
    frames_r[1, :, :, :] = 0
    yy, xx = np.ogrid[:hin, :win]
    frames_r[1][(xx - 100) ** 2 + (yy - 100) ** 2 <= 70 * 70] = 123  # a smaller disc in one right frame
    left, right = torch.from_numpy(frames_l).cuda(), torch.from_numpy(frames_r).cuda()
    # per-eye tuple, fixed radius
    wp = V.SbsWarper((tl, tr), size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius=95.0)
    got = wp(left, right).cpu().numpy()
    for f in range(n):
        want = V.lr_frame((tl, tr), frames_l[f], frames_r[f], size_output=(wout, hout), interpolation=1, radius=95.0)
        assert np.array_equal(got[f], want)
    # auto radius per frame, consumed on the device
    wp = V.SbsWarper(tl, size_input=(hin, win), size_output=(wout, hout), interpolation=2, radius="auto")
    got = wp(left, right).cpu().numpy()
    rad, trans = wp.radius_per_frame(left, right)
    rad = rad.cpu().numpy()
    for f in range(n):
        want_r = max(chain_np.get_radius(frames_l[f]), chain_np.get_radius(frames_r[f]))
        assert rad[f] == want_r
        want = V.lr_frame(tl, frames_l[f], frames_r[f], size_output=(wout, hout), interpolation=2, radius="auto")
        assert np.array_equal(got[f], want), f


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("per_eye", [False, True])
def test_tiled_pipeline_ring_wraparound_and_source_edges(interp, per_eye):
    """Source-edge handling of the tiled TMA kernel (csrc/tiled.cu) on an odd batch: a radius larger than the source
    so that many tiles straddle the source edge (TMA zero fill == BORDER_CONSTANT 0), tiles fully outside, partial
    edge tiles.  NOTE: with so few tiles the launcher gives every CTA ONE frame, so the stage ring is never re-filled
    here -- the long frame loops (ring refills, mbarrier phase flips, > 1 frame per CTA with a per-frame radius) are
    covered by tests/test_gpu_frameloop.py, which forces the frames per CTA.  (Output 208 x 104 is not a
    multiple of the 32 x 32 / 32 x 16 / 32 x 8 tiles; 208 * 3 bytes keeps the right eye's column offset 16-byte aligned,
    which the TMA path needs.)  Shared map (2 views per CTA) and per-eye maps (1 view per CTA)."""
    import torch

    n, hin, win, wout, hout = 23, 192, 224, 208, 104
    ql, qr = V.from_rotation_vector([0.03, -0.02, 0.05]), V.from_rotation_vector([-0.03, 0.02, -0.05])
    mk = lambda q: (V.EquirectangularEncoder() * V.Euclidean3DRotator(q) * V.PolynomialScaler([0, 1, 0.05])  # noqa: E731
                    * V.FisheyeDecoder("equidistant"))
    t = (mk(ql), mk(qr)) if per_eye else mk(ql)
    rng = np.random.default_rng(7)
    ln = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)  # no black surround: edges carry signal
    rn = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    radius = 150.0  # > min(hin, win) / 2: the footprint leaves the source on every side
    wp = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=radius)
    got = wp(left, right).cpu().numpy()

    def omap(q):
        ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q.components).ravel().tolist()),
               ("poly", [0, 1, 0.05]), ("fisheye_dec", "equidistant")]
        return chain_np.get_map(ops, radius=radius, size_input=(hin, win), size_output=(wout, hout))

    ml = omap(ql)
    mr = omap(qr) if per_eye else ml
    for f in range(n):
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp),
                               cv2.remap(rn[f], mr[0], mr[1], interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, per_eye, f, int((got[f] != want).sum()))


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
@pytest.mark.parametrize("per_eye", [False, True])
def test_tiled_per_frame_radius_on_device(interp, per_eye):
    """Per-frame radius consumed on the device by the tiled kernel (vr180_mapsrc_t::radius_dev): every frame has
    its own source rectangle.  Radii: small, larger than the source (edges), negative / x.5 (what get_radius
    really returns) and NaN (no transition found -> all border)."""
    import torch

    hin, win, wout, hout = 192, 224, 160, 96
    radii = [90.0, 150.0, float("nan"), -100.5, 60.25, 111.0, 149.5, -75.0, 33.0, 128.0, 95.5]
    n = len(radii)
    ql, qr = V.from_rotation_vector([0.03, -0.02, 0.05]), V.from_rotation_vector([-0.03, 0.02, -0.05])
    mk = lambda q: (V.EquirectangularEncoder() * V.Euclidean3DRotator(q) * V.PolynomialScaler([0, 1, 0.05])  # noqa: E731
                    * V.FisheyeDecoder("equidistant"))
    t = (mk(ql), mk(qr)) if per_eye else mk(ql)
    rng = np.random.default_rng(11)
    ln = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    wp = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius="auto")
    got = wp(left, right, radius=torch.tensor(radii, dtype=torch.float64, device="cuda")).cpu().numpy()

    def omap(q, r):
        ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q.components).ravel().tolist()),
               ("poly", [0, 1, 0.05]), ("fisheye_dec", "equidistant")]
        return chain_np.get_map(ops, radius=r, size_input=(hin, win), size_output=(wout, hout))

    for f, r in enumerate(radii):
        if r != r:
            assert not got[f].any(), f  # NaN coordinates sample the (zero) border colour everywhere
            continue
        ml = omap(ql, r)
        mr = omap(qr, r) if per_eye else ml
        want = np.concatenate([cv2.remap(ln[f], ml[0], ml[1], interpolation=interp),
                               cv2.remap(rn[f], mr[0], mr[1], interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, per_eye, f, r, int((got[f] != want).sum()))


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_tiled_per_frame_radius_equal_radii_take_fixed_pipeline(interp):
    """A chunk whose frames all carry the same device radius (a static rig) is routed to the fixed-radius pipeline
    inside the per-frame-radius kernel (csrc/tiled.cu, `dynr`): the result must be the fixed-radius result, bit for
    bit, for an odd frame count (two-frame items leave a phantom frame) and for both interpolations."""
    import torch

    hin, win, wout, hout = 192, 224, 160, 96
    n, r = 7, 101.5
    q = V.from_rotation_vector([0.03, -0.02, 0.05])
    t = V.EquirectangularEncoder() * V.Euclidean3DRotator(q) * V.PolynomialScaler([0, 1, 0.05]) * V.FisheyeDecoder("equidistant")
    rng = np.random.default_rng(13)
    ln = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    left, right = torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()
    dyn = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius="auto")
    got = dyn(left, right, radius=torch.full((n,), r, dtype=torch.float64, device="cuda")).cpu().numpy()
    fixed = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=r)
    want = fixed(left, right).cpu().numpy()
    assert np.array_equal(got, want)
    ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q.components).ravel().tolist()),
           ("poly", [0, 1, 0.05]), ("fisheye_dec", "equidistant")]
    xm, ym = chain_np.get_map(ops, radius=r, size_input=(hin, win), size_output=(wout, hout))
    for f in (0, n - 1):
        ref = np.concatenate([cv2.remap(ln[f], xm, ym, interpolation=interp), cv2.remap(rn[f], xm, ym, interpolation=interp)], axis=1)
        assert np.array_equal(got[f], ref), (interp, f)


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_tiled_long_batch_many_ring_wraps(interp):
    """131 frames through one launch with the AUTOMATIC frames-per-CTA choice (6-12 tiles -> one frame per CTA, 131
    frame chunks in grid.z): checks the chunk indexing over a long batch.  The ring-wrap regime itself is forced and
    checked in tests/test_gpu_frameloop.py.  Every frame is compared with cv2.remap on the oracle's maps."""
    import torch

    n, hin, win, wout, hout = 131, 96, 128, 96, 64
    q = V.from_rotation_vector([0.02, 0.03, -0.04])
    t = V.EquirectangularEncoder() * V.Euclidean3DRotator(q) * V.FisheyeDecoder("equidistant")
    rng = np.random.default_rng(17)
    ln = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    got = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=48.0)(
        torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda()).cpu().numpy()
    ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(*q.components).ravel().tolist()), ("fisheye_dec", "equidistant")]
    xm, ym = chain_np.get_map(ops, radius=48.0, size_input=(hin, win), size_output=(wout, hout))
    for f in range(n):
        want = np.concatenate([cv2.remap(ln[f], xm, ym, interpolation=interp), cv2.remap(rn[f], xm, ym, interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, f)


@pytest.mark.parametrize("interp", [1, 2])
def test_tiled_multi_frame_items_with_phantom_frames(interp):
    """Enough tiles (1024 x 512 output) for the launcher to keep >= 8 frames per CTA: bicubic then packs FOUR frames
    per pipeline item, bilinear two.  22 frames -> chunks of 12 and 10 frames: the last item of the launch has two
    phantom frames past the end of the batch (TMA zero-fills their loads and drops their stores)."""
    import torch

    n, hin, win, wout, hout = 22, 320, 320, 1024, 512
    t = V.EquirectangularEncoder() * V.PolynomialScaler([0, 1, 0.03]) * V.FisheyeDecoder("equidistant")
    rng = np.random.default_rng(23)
    ln = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    rn = rng.integers(0, 256, (n, hin, win, 3), dtype=np.uint8)
    out = torch.full((n + 1, hout, 2 * wout, 3), 77, dtype=torch.uint8, device="cuda")  # one guard frame behind the batch
    V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=interp, radius=160.0)(
        torch.from_numpy(ln).cuda(), torch.from_numpy(rn).cuda(), out=out[:n])
    got = out.cpu().numpy()
    assert (got[n] == 77).all(), "a phantom frame was stored behind the batch"
    xm, ym = chain_np.get_map([("equirect_enc", True), ("poly", [0, 1, 0.03]), ("fisheye_dec", "equidistant")], radius=160.0,
                              size_input=(hin, win), size_output=(wout, hout))
    for f in range(n):
        want = np.concatenate([cv2.remap(ln[f], xm, ym, interpolation=interp), cv2.remap(rn[f], xm, ym, interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, f)


def test_tiled_sbs_offset_not_16_byte_aligned_takes_generic_kernel():
    """TMA boxes must start at 16-byte aligned global addresses: an output width whose eye offset (W * 3 bytes) is
    not a multiple of 16 is not eligible for the tiled kernel and must still be exact through the generic one."""
    import torch

    n, hin, win, wout, hout = 3, 160, 160, 200, 64  # 200 * 3 = 600 bytes: 8 mod 16
    t = V.EquirectangularEncoder() * V.FisheyeDecoder("equidistant")
    left = torch.from_numpy(_frames(n, hin, win, 0)).cuda()
    right = torch.from_numpy(_frames(n, hin, win, 100)).cuda()
    got = V.SbsWarper(t, size_input=(hin, win), size_output=(wout, hout), interpolation=1, radius=80.0)(left, right).cpu().numpy()
    xm, ym = chain_np.get_map([("equirect_enc", True), ("fisheye_dec", "equidistant")], radius=80.0, size_input=(hin, win),
                              size_output=(wout, hout))
    ln, rn = left.cpu().numpy(), right.cpu().numpy()
    for f in range(n):
        want = np.concatenate([cv2.remap(ln[f], xm, ym, interpolation=1), cv2.remap(rn[f], xm, ym, interpolation=1)], axis=1)
        assert np.array_equal(got[f], want), f


# ---------------------------------------------------------------------------------------------------------
# (6) full BASELINE sizes through size-independent properties
# ---------------------------------------------------------------------------------------------------------
def test_fullsize_8k_pair_properties():
    """cfg3-sized pair (2 x 4096^2 -> 8192 x 4096): (a) identity-like check against cv2 on the GPU-built maps,
    (b) both eyes land in their halves, (c) batch invariance (frame f of a batch == the same frame alone)."""
    import torch

    n = 4096
    q = V.from_rotation_vector([0.01, 0.02, -0.015])
    t = V.EquirectangularEncoder() * V.Euclidean3DRotator(q) * V.PolynomialScaler([0, 1, -0.02, 0.003]) * V.FisheyeDecoder("equidistant")
    fl, fr = disc_frame(n, n, 0), disc_frame(n, n, 1)
    left = torch.from_numpy(np.stack([fl, fr])).cuda()
    right = torch.from_numpy(np.stack([fr, fl])).cuda()
    wp = V.SbsWarper(t, size_input=(n, n), size_output=(n, n), interpolation=1, radius=n / 2)
    out = wp(left, right)
    assert out.shape == (2, n, 2 * n, 3)
    # (c) swapping the eyes swaps the halves, bit for bit
    assert torch.equal(out[0, :, :n], out[1, :, n:]) and torch.equal(out[0, :, n:], out[1, :, :n])
    # (a) fused analytic output == cv2.remap driven by the GPU-built float32 maps (pixel exactness given the maps)
    maps = wp.maps()[0].cpu().numpy()
    want = cv2.remap(fl, maps[0], maps[1], interpolation=1)
    assert np.array_equal(out[0, :, :n].cpu().numpy(), want)
    # checksum of checksums: LUT and fixed-LUT sources give the same frame
    for src in ("lut", "lut_fixed"):
        other = V.SbsWarper(t, size_input=(n, n), size_output=(n, n), interpolation=1, radius=n / 2, map_source=src)
        assert torch.equal(other(left[:1], right[:1])[0], out[0])


def test_cfg1_reference_cli_run_matches_golden(golden_cfg1):
    """BASELINE.json configs[0]: `v1c lr docs/_static/test.jpg docs/_static/test.jpg` with the default PolynomialScaler
    chain, INTER_LINEAR, CLI default size 4096 x 4096 per eye (cli.py:117-380, remapper.py:448-456: one file split into
    two PORTRAIT 2048 x 1024 views -> get_radius scans the centre COLUMN -> 877.5 -> 4x upsampling into an
    8192 x 4096 SBS frame).  Compared with 200 000 sparse samples + 8 full rows of the unmodified reference's output;
    budget: 4 pixels (one-ulp float32 map flips of another libm), in practice 0 -> the sha256 of the whole frame."""
    import hashlib

    z, meta, jpg = golden_cfg1
    img = cv2.imread(str(jpg))
    if hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() != meta["input_sha256"]:
        pytest.skip("this cv2 build decodes the JPEG differently from the one that generated the fixture")
    t = eval(meta["expr"], NS)  # noqa: S307
    halves = [img[:, : img.shape[1] // 2], img[:, img.shape[1] // 2:]]
    assert V.get_radius_smart("auto", halves) == meta["radius"] == 877.5
    out = V.lr_frame(t, str(jpg), str(jpg), size_output=(4096, 4096), interpolation=1, radius="auto")
    assert out.shape == (4096, 8192, 3)
    idx = z["idx"]
    bad = int((out[idx[:, 0], idx[:, 1]] != z["px"]).any(axis=1).sum()) + int((out[z["rows"]] != z["row_px"]).any(axis=2).sum())
    assert bad <= 4, bad
    if bad == 0:
        digest = hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest()
        print("cfg1 sha256", digest, "reference", meta["output_sha256"])


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_packed_lut_tiles_with_unpackable_tiles(interp):
    """vr180_pack_lut_tiles on maps that contain NaN / huge coordinates, a footprint wider than 255 px and partial edge
    tiles: those tiles are flagged and read the float32 maps; every pixel must equal cv2.remap on the same maps."""
    import torch

    rng = np.random.default_rng(31 + interp)
    hin, win, wout, hout = 300, 420, 208, 104  # 208 x 104: partial bottom tiles for every mode
    src = rng.integers(0, 256, (2, hin, win, 3), dtype=np.uint8)
    yy, xx = np.mgrid[:hout, :wout].astype(np.float32)
    xm = xx * 1.3 + 20 + 0.37 * np.sin(yy / 7).astype(np.float32)
    ym = yy * 1.7 + 15 + 0.41 * np.cos(xx / 9).astype(np.float32)
    xm[0:32, 0:32] = np.nan               # a tile of NaN coordinates
    xm[40, 40] = 1e9                      # saturated coordinate
    xm[32:64, 64:96] = (xx[32:64, 64:96] - 64) * 12.5   # footprint 400 px wide: cannot be packed (and not staged)
    ym[70, 100] = -np.inf

    class Maps(V.TransformerBase):  # an opaque transformer that returns these maps: the plan goes through host_maps
        def transform(self, x, y, **kw):
            return xm.astype(np.float64), ym.astype(np.float64)

        def inverse_transform(self, x, y, **kw):
            raise NotImplementedError

    left = torch.from_numpy(src).cuda()
    wp = V.SbsWarper(V.MultiTransformer(transformers=[Maps()]), size_input=(hin, win), size_output=(wout, hout),
                     interpolation=interp, radius=1.0, map_source="lut_packed")
    wp._maps = torch.from_numpy(np.stack([xm, ym])[None]).cuda()  # bypass Normalize / Denormalize: these ARE the maps
    got = wp(left, left.flip(0)).cpu().numpy()
    for f in range(2):
        want = np.concatenate([cv2.remap(src[f], xm, ym, interpolation=interp), cv2.remap(src[1 - f], xm, ym, interpolation=interp)], axis=1)
        assert np.array_equal(got[f], want), (interp, f, int((got[f] != want).sum()))


def test_apply_lr_merge_matches_reference_golden(golden_merge_match, golden_apply, golden_maps, tmp_path):
    """apply_lr(merge=True) (remapper.py:485-516): device anaglyph kernel + host labels == the reference's PNG, with
    and without labels (font scale rows // 1000), shared transformer and per-eye tuple with radius="auto"."""
    g = golden_merge_match
    _, meta = golden_maps
    card = golden_apply["card"]
    left, right = card[:, :128], card[:, 128:]
    t = eval(meta["cases"]["rot_poly"]["expr"], NS)  # noqa: S307
    out = tmp_path / "m.png"
    for name, size in (("small", (96, 80)), ("labels", (512, 1024))):
        V.apply_lr(t, left_path=left, right_path=right, out_path=out, size_output=size, interpolation=1, radius="max",
                   merge=True)
        got = cv2.imread(str(out))
        assert got.shape == g[f"merge/{name}"].shape and np.array_equal(got, g[f"merge/{name}"]), name
    tr = eval(meta["cases"]["rot_nonunit"]["expr"], NS)  # noqa: S307
    V.apply_lr((t, tr), left_path=left, right_path=right, out_path=out, size_output=(96, 80), interpolation=2,
               radius="auto", merge=True)
    assert np.array_equal(cv2.imread(str(out)), g["merge/tuple_auto"])
    # the device kernel alone against the oracle restatement on random eyes (incl. exact .5 ties of the float64 value)
    rng = np.random.default_rng(3)
    a, b = rng.integers(0, 256, (2, 70, 90, 3), dtype=np.uint8)
    from vr180_convert_b200.remapper import _merge_device

    assert np.array_equal(_merge_device(a, b), remap_np.anaglyph_u8(a, b))


def test_match_lr_matches_reference_golden(golden_merge_match, golden_apply):
    """match_lr (remapper.py:251-321): vr180_transform_points (float64) vs the reference's float32 NumPy chain."""
    g = golden_merge_match
    card = golden_apply["card"]
    imgs = [np.ascontiguousarray(card[:, :128]), np.ascontiguousarray(card[:, 128:])]
    pts_l, pts_r = g["match/pts_l"], g["match/pts_r"]
    dec = V.FisheyeDecoder("equidistant")
    for rname, radius in (("max", "max"), ("auto", "auto"), ("60.5", 60.5)):
        vl, vr = V.match_lr(dec, pts_l, pts_r, imgs, radius=radius)
        np.testing.assert_allclose(vl, g[f"match/single/{rname}/vl"], rtol=0, atol=3e-6, equal_nan=True)
        np.testing.assert_allclose(vr, g[f"match/single/{rname}/vr"], rtol=0, atol=3e-6, equal_nan=True)
    dec2 = (V.FisheyeDecoder("stereographic"), V.ZoomTransformer(1.1) * V.FisheyeDecoder("equisolid"))
    vl, vr = V.match_lr(dec2, pts_l, pts_r, imgs, radius=70.0)
    np.testing.assert_allclose(vl, g["match/tuple/70.0/vl"], rtol=0, atol=3e-6, equal_nan=True)
    np.testing.assert_allclose(vr, g["match/tuple/70.0/vr"], rtol=0, atol=3e-6, equal_nan=True)
    with pytest.raises(ValueError):
        V.match_lr(dec, pts_l[:3], pts_r, imgs, radius=50.0)


@pytest.mark.parametrize("interp", [1, 4])
def test_host_pipeline_output_width_not_16_byte_aligned(interp):
    """--size 1000x1000-style outputs (W * 3 not a multiple of 16 bytes): the host pipeline puts the right eye at a
    16-byte aligned column of its device frame (so the launch stays on the TMA-tiled kernel) and downloads the eyes as
    two column segments into the dense host frame; pinned and pageable destinations, SBS and merged output."""
    hin, win, wout, hout = 160, 160, 100, 72  # 300 bytes per eye row: 12 mod 16
    t = V.EquirectangularEncoder() * V.PolynomialScaler([0, 1, 0.03]) * V.FisheyeDecoder("equidistant")
    lefts = [disc_frame(hin, win, seed=i) for i in range(5)]
    rights = [disc_frame(hin, win, seed=50 + i) for i in range(5)]
    xm, ym = chain_np.get_map([("equirect_enc", True), ("poly", [0, 1, 0.03]), ("fisheye_dec", "equidistant")], radius=80.0,
                              size_input=(hin, win), size_output=(wout, hout))
    got = V.lr_frames(t, lefts, rights, size_output=(wout, hout), interpolation=interp, radius=80.0)
    import os
    os.environ["VR180_PINNED_OUTPUTS"] = "0"  # pageable result arrays: the staged (drain thread) download path
    try:
        got_pageable = V.lr_frames(t, lefts, rights, size_output=(wout, hout), interpolation=interp, radius=80.0)
    finally:
        del os.environ["VR180_PINNED_OUTPUTS"]
    for f in range(5):
        eyes = [cv2.remap(img, xm, ym, interpolation=interp) for img in (lefts[f], rights[f])]
        want = np.concatenate(eyes, axis=1)
        assert got[f].shape == (hout, 2 * wout, 3) and np.array_equal(got[f], want), (interp, f)
        assert np.array_equal(got_pageable[f], want), (interp, f, "pageable")
        merged = V.lr_frame(t, lefts[f], rights[f], size_output=(wout, hout), interpolation=interp, radius=80.0, merge=True)
        assert np.array_equal(merged, remap_np.anaglyph_u8(eyes[0], eyes[1])), (interp, f, "merge")


def test_nvjpeg_codec_path(tmp_path, golden_apply):
    """Opt-in device codec (include/vr180_b200.h section 6): nvJPEG decode agrees with cv.imread up to the decoder
    differences the header documents (NOT bit-exact: stated tolerance mean |diff| < 1.5 grey levels, < 1 % of the
    samples off by more than 8), and the JPEG -> warp -> JPEG form of apply_lr on the device produces the same picture
    as the cv2 path up to codec noise (PSNR > 30 dB).  The default codec stays cv2 (bit-exact tests above)."""
    import torch

    from vr180_convert_b200 import codec

    if not codec.available():
        pytest.skip("libnvjpeg could not be loaded")
    # a smooth picture (blurred noise + a disc): at sharp colour edges the two decoders differ by design -- libjpeg-turbo
    # interpolates 4:2:0 chroma ("fancy upsampling"), nvJPEG replicates it -- which is why the codec is opt-in
    rng = np.random.default_rng(0)
    card = cv2.GaussianBlur(rng.integers(0, 256, (512, 512, 3), dtype=np.uint8), (0, 0), 6)
    card = cv2.normalize(card, None, 0, 255, cv2.NORM_MINMAX)
    yy, xx = np.ogrid[:512, :512]
    outside = ((xx - 128) ** 2 + (yy - 256) ** 2 > 120 ** 2) & ((xx - 384) ** 2 + (yy - 256) ** 2 > 120 ** 2)
    card[outside] = 0
    src = tmp_path / "sbs.jpg"
    cv2.imwrite(str(src), card)
    want = cv2.imread(str(src))
    got = codec.decode_jpeg_device(src).cpu().numpy()
    assert got.shape == want.shape
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print("nvjpeg vs cv2 decode: mean |diff|", float(d.mean()), "max", int(d.max()), "> 8:", float((d > 8).mean()))
    assert d.mean() < 2.0 and (d > 16).mean() < 0.02, (float(d.mean()), int(d.max()))
    # 4:4:4 file: no chroma upsampling involved, only the IDCT / colour conversion roundings differ
    src444 = tmp_path / "c444.jpg"
    cv2.imwrite(str(src444), card, [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444])
    d444 = np.abs(codec.decode_jpeg_device(src444).cpu().numpy().astype(np.int32) - cv2.imread(str(src444)).astype(np.int32))
    print("4:4:4: mean |diff|", float(d444.mean()), "max", int(d444.max()))
    assert d444.mean() < 1.0 and d444.max() <= 8, (float(d444.mean()), int(d444.max()))
    # encode: round trip through cv2's decoder
    enc = codec.encode_jpeg_device(torch.from_numpy(want).cuda())
    back = cv2.imdecode(np.frombuffer(enc, np.uint8), cv2.IMREAD_COLOR)
    mse = np.mean((back.astype(np.float64) - want.astype(np.float64)) ** 2)
    assert back.shape == want.shape and 10 * np.log10(255.0 ** 2 / mse) > 30
    # whole apply_lr on the device vs the cv2 path
    t = V.EquirectangularEncoder() * V.PolynomialScaler([0, 1, 0.02]) * V.FisheyeDecoder("equidistant")
    out_ref, out_dev = tmp_path / "ref.jpg", tmp_path / "dev.jpg"
    V.apply_lr(t, left_path=src, right_path=src, out_path=out_ref, size_output=(256, 256), interpolation=1, radius="max")
    V.set_codec("nvjpeg")
    try:
        V.apply_lr(t, left_path=src, right_path=src, out_path=out_dev, size_output=(256, 256), interpolation=1, radius="max")
        V.apply_lr(t, left_path=src, right_path=src, out_path=tmp_path / "dev.png", size_output=(256, 256), interpolation=1,
                   radius="auto", merge=True)
    finally:
        V.set_codec("cv2")
    a, b = cv2.imread(str(out_ref)).astype(np.float64), cv2.imread(str(out_dev)).astype(np.float64)
    assert a.shape == b.shape == (256, 512, 3)
    assert 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2)) > 30
    assert cv2.imread(str(tmp_path / "dev.png")).shape == (256, 256, 3)
