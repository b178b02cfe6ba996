"""CPU tests (-m "not gpu"): the oracle restatements against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py) and against the installed cv2 (the reference's own dependency)."""
from __future__ import annotations

import cv2
import numpy as np
import pytest

from oracle import chain_np, remap_np
from tests.conftest import disc_frame


def _ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(2**31) - ai, ai)
    bi = np.where(bi < 0, -(2**31) - bi, bi)
    return np.abs(ai - bi)


def test_chain_oracle_matches_reference_maps(golden_maps):
    """Same host, same NumPy: the step-by-step restatement must reproduce the reference float32 maps
    bit-for-bit (NaN positions included)."""
    z, meta = golden_maps
    for name, case in meta["cases"].items():
        for si, (size_out, size_in, radius) in enumerate(meta["shapes"]):
            xm, ym = chain_np.get_map(case["ops"], radius=radius, size_input=tuple(size_in),
                                      size_output=tuple(size_out))
            for got, key in ((xm, "x"), (ym, "y")):
                want = z[f"{name}/{si}/{key}"]
                assert got.shape == want.shape and got.dtype == np.float32
                assert np.array_equal(np.isnan(got), np.isnan(want)), name
                ok = ~np.isnan(want)
                # identical on the generating host; allow 1 float32 ulp for a different libm / SIMD width
                assert _ulp_diff(got[ok], want[ok]).max(initial=0) <= 1, (name, si, key)


def test_chain_oracle_fullsize_samples(golden_fullsize):
    z = golden_fullsize
    ops = {"base": [("equirect_enc", True), ("fisheye_dec", "equidistant")]}
    n = 2048
    xm, ym = chain_np.get_map(ops["base"], radius=n / 2, size_input=(n, n), size_output=(n, n))
    idx = z[f"base/{n}/idx"]
    assert _ulp_diff(xm[idx[:, 0], idx[:, 1]], z[f"base/{n}/x"]).max() <= 1
    assert _ulp_diff(ym[idx[:, 0], idx[:, 1]], z[f"base/{n}/y"]).max() <= 1
    assert _ulp_diff(xm[n // 3], z[f"base/{n}/row"][0]).max() <= 1
    assert _ulp_diff(ym[:, n // 5], z[f"base/{n}/col"][1]).max() <= 1


def test_survey_known_answers():
    """SURVEY.md Appendix E known-answer values (float32 bit patterns)."""
    xm, ym = chain_np.get_map([("equirect_enc", True), ("fisheye_dec", "equidistant")], radius=128.0,
                              size_input=(256, 256), size_output=(256, 256))
    assert (xm[0, 0], ym[0, 0]) == (128.0, 0.0)
    assert (xm[128, 0], ym[128, 0]) == (0.0, 128.0)
    assert (xm[128, 128], ym[128, 128]) == (128.0, 128.0)
    assert xm[37, 201].view(np.uint32) == 0x432585C3 and ym[37, 201].view(np.uint32) == 0x41EC3D23
    assert xm[255, 255].view(np.uint32) == 0x4301920C and ym[255, 255].view(np.uint32) == 0x437FFA64
    rmat = [0.9996874316381736, 0.01510030258717884, 0.01992535787502085, -0.014900258835609935,
            0.9998374644518503, -0.010150219954606273, -0.020075390688697525, 0.009850154327252914,
            0.9997499453105388]
    xm, ym = chain_np.get_map([("equirect_enc", True), ("rot3", rmat), ("poly", [0, 1, -0.02, 0.003]),
                               ("fisheye_dec", "equidistant")], radius=128.0, size_input=(256, 256),
                              size_output=(256, 256))
    assert xm[0, 0].view(np.uint32) == 0x42FC3427 and ym[0, 0].view(np.uint32) == 0x40145B3C
    assert xm[37, 201].view(np.uint32) == 0x4324244F and ym[37, 201].view(np.uint32) == 0x41EFE0B8
    assert xm[255, 255].view(np.uint32) == 0x430365FD and ym[255, 255].view(np.uint32) == 0x437C1DFE


def test_quaternion_matrix_matches_scipy():
    from scipy.spatial.transform import Rotation

    for q in ((0.9999093510664558, 0.00500054686470522, 0.01000109372941044, -0.00750082029705783),
              (1.3, 0.02, -0.05, 0.04)):
        m = chain_np.quat_to_matrix(*q)
        w, x, y, z = q
        want = Rotation.from_quat([x, y, z, w]).as_matrix()  # scipy is scalar-last and normalises
        assert np.abs(m - want).max() < 1e-15


def test_remap_oracle_matches_golden(golden_remap):
    g = golden_remap
    src, xm, ym = g["src"], g["xmap"], g["ymap"]
    for interp in (0, 1, 2, 4):
        for bm in (0, 1, 2, 3, 4):
            for bi, bv in enumerate((0, 7, (3, 200, 90))):
                got = remap_np.remap(src, xm, ym, interp, bm, bv)
                assert np.array_equal(got, g[f"out/{interp}/{bm}/{bi}"]), (interp, bm, bi)


@pytest.mark.parametrize("interp", [0, 1, 2, 4])
def test_remap_oracle_matches_live_cv2(interp):
    """The reference's actual third-party call, live (guards against cv2 version drift on the box)."""
    rng = np.random.default_rng(interp)
    src = disc_frame(96, 128, seed=3)
    xm = (rng.random((70, 90)) * 140 - 6).astype(np.float32)
    ym = (rng.random((70, 90)) * 110 - 7).astype(np.float32)
    xm[0, :3] = [np.nan, np.inf, -1e9]
    for bm in (0, 1, 2, 3, 4):
        want = cv2.remap(src, xm, ym, interpolation=interp, borderMode=bm, borderValue=0)
        assert np.array_equal(remap_np.remap(src, xm, ym, interp, bm, 0), want)
    view = src[:, 10:90]  # row-strided input view, as remapper.py:455-456 produces
    assert np.array_equal(remap_np.remap(view, xm, ym, interp), cv2.remap(view, xm, ym, interpolation=interp))


def test_weight_tables_sum():
    for interp in (1, 2, 4):
        t = remap_np.weight_table(interp).astype(np.int64)
        assert (t.reshape(1024, -1).sum(axis=1) == 32768).all()


def test_get_radius_oracle(golden_radius):
    for key, want in golden_radius.items():
        if "x" in key and key[0].isdigit():
            h, w, r = map(int, key.split("x"))
            yy, xx = np.mgrid[:h, :w]
            img = np.where(((xx - w // 2) ** 2 + (yy - h // 2) ** 2 <= r * r)[..., None], 200, 0).astype(np.uint8)
            img = np.repeat(img, 3, axis=2)
            assert chain_np.get_radius(img) == want
            pos, neg = chain_np.get_radius_transitions(img)
            assert (neg - pos) / 2 == want
    for name in ("black64", "white64"):
        img = np.zeros((64, 64, 3), np.uint8) if name == "black64" else np.full((64, 64, 3), 255, np.uint8)
        with pytest.raises(IndexError):
            chain_np.get_radius(img)
        assert -1 in chain_np.get_radius_transitions(img)


def test_cfg1_reference_cli_run(golden_cfg1):
    """BASELINE.json configs[0] / SURVEY.md Appendix E: `v1c lr test.jpg test.jpg` with PolynomialScaler(), INTER_LINEAR,
    4096 x 4096 per eye (cli.py:117-380 -> remapper.py:406-520).  The oracle (get_radius + chain restatement + cv2.remap
    per half + concatenate) reproduces the reference's logged radius and its output samples; on the host that generated
    the fixture the whole 100 MB frame hashes to the recorded sha256."""
    import hashlib

    z, meta, jpg = golden_cfg1
    img = cv2.imread(str(jpg))
    if hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() != meta["input_sha256"]:
        pytest.skip("this cv2 build decodes the JPEG differently from the one that generated the fixture")
    halves = [img[:, : img.shape[1] // 2], img[:, img.shape[1] // 2:]]  # remapper.py:448-456: views, portrait 2048 x 1024
    radius = max(chain_np.get_radius(h) for h in halves)  # remapper.py:82-84
    assert radius == meta["radius"] == 877.5
    ops = [("equirect_enc", True), ("poly", [0, 1]), ("fisheye_dec", "equidistant")]
    xm, ym = chain_np.get_map(ops, radius=radius, size_input=halves[0].shape[:2], size_output=(4096, 4096))
    out = np.concatenate([cv2.remap(h, xm, ym, interpolation=cv2.INTER_LINEAR) for h in halves], axis=1)
    assert out.shape == (4096, 8192, 3)
    idx = z["idx"]
    bad = int((out[idx[:, 0], idx[:, 1]] != z["px"]).any(axis=1).sum()) + int((out[z["rows"]] != z["row_px"]).any(axis=2).sum())
    assert bad <= 4, bad  # another libm may flip a vanishing number of float32 map values by one ulp
    if bad == 0 and np.__version__ == meta["numpy"]:
        assert hashlib.sha256(out.tobytes()).hexdigest() == meta["output_sha256"]


def test_anaglyph_oracle_matches_reference_merge(golden_merge_match, golden_apply):
    """remap_np.anaglyph_u8 (+ LINE_8 labels) == the PNG the reference's apply_lr(merge=True) writes."""
    g = golden_merge_match
    card = golden_apply["card"]
    left, right = card[:, :128], card[:, 128:]
    ops = [("equirect_enc", True), ("rot3", chain_np.quat_to_matrix(0.9999093510664558, 0.00500054686470522, 0.01000109372941044,
                                                                      -0.00750082029705783).ravel().tolist()),
           ("poly", [0, 1, -0.02, 0.003]), ("fisheye_dec", "equidistant")]
    for name, size in (("small", (96, 80)), ("labels", (512, 1024))):
        xm, ym = chain_np.get_map(ops, radius=64.0, size_input=(256, 128), size_output=size)  # radius "max" = min(h/2, w/2)
        eyes = [cv2.remap(np.ascontiguousarray(e), xm, ym, interpolation=cv2.INTER_LINEAR) for e in (left, right)]
        got = remap_np.anaglyph_u8(eyes[0], eyes[1])
        colors = [(0, 128, 255), (255, 128, 0)]
        got = np.ascontiguousarray(got)
        cv2.putText(got, "L", (0, len(got[1]) // 10), cv2.FONT_HERSHEY_SIMPLEX, len(got) // 1000, colors[0], 2, cv2.LINE_8)
        cv2.putText(got, "R", (len(got[1]) // 2, len(got[0]) // 10), cv2.FONT_HERSHEY_SIMPLEX, len(got) // 1000, colors[1],
                    2, cv2.LINE_8)
        assert np.array_equal(got, g[f"merge/{name}"]), name
