"""ORACLE (test infrastructure, NOT product code) -- integer restatement of `cv2.remap` for uint8 images,
the third-party call the reference makes at /root/reference/src/vr180_convert/remapper.py:388-398.

OpenCV's source is not under /root/reference (opencv-python pinned 4.10.0.82 in poetry.lock:1014-1015;
4.13.0.92 is installed in this image).  This file restates the published algorithm of
`cv::remap` (modules/imgproc/src/imgwarp.cpp: INTER_BITS=5, INTER_TAB_SIZE=32,
INTER_REMAP_COEF_BITS=15) in vectorised NumPy and is pinned two ways:

  * live, on every test run: tests/test_oracle_golden.py compares it with the installed `cv2.remap`
    (every interpolation x border mode, adversarial maps incl. NaN/inf/huge values);
  * frozen: tests/golden/remap_*.npz hold `cv2.remap` outputs generated in the build container.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

INTER_NEAREST, INTER_LINEAR, INTER_CUBIC, INTER_LANCZOS4 = 0, 1, 2, 4
BORDER_CONSTANT, BORDER_REPLICATE, BORDER_REFLECT, BORDER_WRAP, BORDER_REFLECT_101 = 0, 1, 2, 3, 4

INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS
_INT_MIN = -(2**31)


def cv_round_f32(v: np.ndarray) -> np.ndarray:
    """cvRound on float32 lanes = x86 `cvtps2dq`: round-half-even, NaN/inf/|v|>=2^31 -> INT_MIN."""
    v = np.asarray(v, dtype=np.float32)
    with np.errstate(invalid="ignore"):
        r = np.rint(v.astype(np.float64))
    bad = ~np.isfinite(r) | (r >= 2.0**31) | (r < -(2.0**31))
    return np.where(bad, float(_INT_MIN), r).astype(np.int64)


def _sat16(v):
    return np.clip(v, -32768, 32767)


def quantise(xmap: np.ndarray, ymap: np.ndarray):
    """Map element -> (ix, iy, ax, ay) for LINEAR/CUBIC/LANCZOS4: float32 multiply by 32, cvRound,
    low 5 bits = fraction index, arithmetic >>5 saturated to int16 = integer tap origin."""
    sx = cv_round_f32(np.asarray(xmap, np.float32) * np.float32(INTER_TAB_SIZE))
    sy = cv_round_f32(np.asarray(ymap, np.float32) * np.float32(INTER_TAB_SIZE))
    return _sat16(sx >> INTER_BITS), _sat16(sy >> INTER_BITS), sx & 31, sy & 31


def border_index(p: np.ndarray, n: int, mode: int) -> np.ndarray:
    """cv::borderInterpolate for the non-constant modes (valid for any integer p)."""
    p = p.astype(np.int64)
    if mode == BORDER_REPLICATE:
        return np.clip(p, 0, n - 1)
    if mode == BORDER_WRAP:
        return np.mod(p, n)
    if mode in (BORDER_REFLECT, BORDER_REFLECT_101):
        if n == 1:
            return np.zeros_like(p)
        # REFLECT      fedcba|abcdefgh|hgfedcb : period 2n,   mirror index 2n-1-q
        # REFLECT_101   gfedcb|abcdefgh|gfedcba : period 2n-2, mirror index 2n-2-q
        period = 2 * n if mode == BORDER_REFLECT else 2 * n - 2
        q = np.mod(p, period)
        return np.where(q < n, q, period - (1 if mode == BORDER_REFLECT else 0) - q)
    raise ValueError(f"unsupported border mode {mode}")


def _fetch(src, ys, xs, mode, cval):
    """src[ys, xs, :] with border handling; ys/xs int64 arrays of identical shape -> (..., C) int64."""
    h, w = src.shape[:2]
    if mode == BORDER_CONSTANT:
        inside = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h)
        v = src[np.clip(ys, 0, h - 1), np.clip(xs, 0, w - 1)].astype(np.int64)
        return np.where(inside[..., None], v, np.asarray(cval, np.int64))
    return src[border_index(ys, h, mode), border_index(xs, w, mode)].astype(np.int64)


# ---------------------------------------------------------------------------------------------------------
# weight tables (initInterTab2D)
# ---------------------------------------------------------------------------------------------------------
def _coef_cubic(t: np.float32) -> np.ndarray:
    f = np.float32
    a = f(-0.75)
    c = np.empty(4, np.float32)
    c[0] = ((a * (t + f(1)) - f(5) * a) * (t + f(1)) + f(8) * a) * (t + f(1)) - f(4) * a
    c[1] = ((a + f(2)) * t - (a + f(3))) * t * t + f(1)
    c[2] = ((a + f(2)) * (f(1) - t) - (a + f(3))) * (f(1) - t) * (f(1) - t) + f(1)
    c[3] = f(1) - c[0] - c[1] - c[2]
    return c


def _coef_lanczos4(t: np.float32) -> np.ndarray:
    c = np.zeros(8, np.float32)
    if t < np.finfo(np.float32).eps:
        c[3] = 1.0
        return c
    s45 = 0.70710678118654752440084436210485
    cs = [[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]]
    x = float(t)
    y0 = -(x + 3) * np.pi * 0.25
    s0, c0 = np.sin(y0), np.cos(y0)
    total = np.float32(0)
    for i in range(8):
        y = -(x + 3 - i) * np.pi * 0.25
        c[i] = np.float32((cs[i][0] * s0 + cs[i][1] * c0) / (y * y))
        total = np.float32(total + c[i])
    inv = np.float32(np.float32(1) / total)
    return (c * inv).astype(np.float32)


_TAB_CACHE: dict[int, np.ndarray] = {}


def weight_table(interpolation: int) -> np.ndarray:
    """int16 table [ay*32+ax][ky][kx]; rows of k*k weights summing to exactly 32768."""
    if interpolation in _TAB_CACHE:
        return _TAB_CACHE[interpolation]
    if interpolation == INTER_LINEAR:
        k = 2
        one = [np.array([np.float32(1) - np.float32(i) * np.float32(1 / 32), np.float32(i) * np.float32(1 / 32)],
                        np.float32) for i in range(32)]
    elif interpolation == INTER_CUBIC:
        k = 4
        one = [_coef_cubic(np.float32(i) * np.float32(1 / 32)) for i in range(32)]
    elif interpolation == INTER_LANCZOS4:
        k = 8
        one = [_coef_lanczos4(np.float32(i) * np.float32(1 / 32)) for i in range(32)]
    else:
        raise ValueError(interpolation)
    tab = np.zeros((1024, k, k), np.int16)
    k2 = k // 2
    for ay in range(32):
        for ax in range(32):
            v = (one[ay][:, None] * one[ax][None, :]).astype(np.float32)
            it = np.clip(np.rint((v * np.float32(COEF_SCALE)).astype(np.float32)), -32768, 32767).astype(np.int64)
            diff = int(it.sum()) - COEF_SCALE
            if diff != 0 and k == 2:
                # LINEAR: only (ax,ay)=(0,0) saturates (32768 -> 32767); OpenCV ends with {32767,0,0,1}, which
                # cannot change a uint8 result (remap() uses the closed 10-bit form for LINEAR anyway)
                it[1, 1] -= diff
            elif diff != 0:
                mk1 = mk2 = Mk1 = Mk2 = k2
                for k1 in range(k2, k2 + 2):
                    for kk in range(k2, k2 + 2):
                        if it[k1, kk] < it[mk1, mk2]:
                            mk1, mk2 = k1, kk
                        elif it[k1, kk] > it[Mk1, Mk2]:
                            Mk1, Mk2 = k1, kk
                if diff < 0:
                    it[Mk1, Mk2] -= diff
                else:
                    it[mk1, mk2] -= diff
            tab[ay * 32 + ax] = it
    _TAB_CACHE[interpolation] = tab
    return tab


# ---------------------------------------------------------------------------------------------------------
# remap
# ---------------------------------------------------------------------------------------------------------
def _border_scalar(value, channels: int):
    """cv2 turns a Python scalar into Scalar(v,0,0,0): only channel 0 is filled."""
    if np.isscalar(value):
        v = [0] * max(channels, 4)
        v[0] = value
    else:
        v = list(value) + [0] * 4
    return np.clip(np.rint(np.asarray(v[:channels], np.float64)), 0, 255).astype(np.int64)


def remap(src: np.ndarray, xmap: np.ndarray, ymap: np.ndarray, interpolation: int = INTER_LINEAR,
          border_mode: int = BORDER_CONSTANT, border_value=0) -> np.ndarray:
    """Bit-exact restatement of cv2.remap(src uint8 HxWxC, float32 maps, ...)."""
    if src.dtype != np.uint8:
        raise TypeError("oracle restates the uint8 path only")
    squeeze = src.ndim == 2
    if squeeze:
        src = src[..., None]
    c = src.shape[2]
    cval = _border_scalar(border_value, c)
    xmap = np.asarray(xmap, np.float32)
    ymap = np.asarray(ymap, np.float32)

    if interpolation == INTER_NEAREST:
        ix = _sat16(cv_round_f32(xmap))
        iy = _sat16(cv_round_f32(ymap))
        out = _fetch(src, iy, ix, border_mode, cval)
    else:
        ix, iy, ax, ay = quantise(xmap, ymap)
        if interpolation == INTER_LINEAR:
            # (w00 p00 + w01 p01 + w10 p10 + w11 p11 + 512) >> 10, weights (32-ax)(32-ay) ... (sum 1024);
            # identical to OpenCV's 15-bit table form for uint8.
            acc = np.zeros(ix.shape + (c,), np.int64)
            for dy, wy in ((0, 32 - ay), (1, ay)):
                for dx, wx in ((0, 32 - ax), (1, ax)):
                    acc += (wy * wx)[..., None] * _fetch(src, iy + dy, ix + dx, border_mode, cval)
            out = (acc + 512) >> 10
        else:
            k = 4 if interpolation == INTER_CUBIC else 8
            tab = weight_table(interpolation).astype(np.int64)[ay * 32 + ax]  # (..., k, k)
            acc = np.zeros(ix.shape + (c,), np.int64)
            off = k // 2 - 1
            for ky in range(k):
                for kx in range(k):
                    acc += tab[..., ky, kx][..., None] * _fetch(src, iy + ky - off, ix + kx - off, border_mode, cval)
            out = np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255)
    out = out.astype(np.uint8)
    return out[..., 0] if squeeze else out


def apply_lr_sbs(src_l, src_r, maps_l, maps_r, interpolation=INTER_LINEAR, border_mode=BORDER_CONSTANT,
                 border_value=0):
    """remapper.py:388-398 per eye + np.concatenate(axis=1) (:518)."""
    left = remap(src_l, maps_l[0], maps_l[1], interpolation, border_mode, border_value)
    right = remap(src_r, maps_r[0], maps_r[1], interpolation, border_mode, border_value)
    return np.concatenate([left, right], axis=1)


def anaglyph_u8(left: np.ndarray, right: np.ndarray) -> np.ndarray:
    """apply_lr(merge=True) without the text labels (/root/reference/src/vr180_convert/remapper.py:485-498), followed
    by the float64 -> uint8 conversion cv.imwrite applies to the reference's float64 image (round half to even,
    saturate; verified against cv2.imwrite + imread)."""
    colors = [(0, 128, 255), (255, 128, 0)]
    combine = np.mean(left, axis=-1)[..., None] * np.array(colors[0]).reshape([1] * (left.ndim - 1) + [3]) + (
        np.mean(right, axis=-1)[..., None] * np.array(colors[1]).reshape([1] * (right.ndim - 1) + [3]))
    combine /= 255
    return np.clip(np.rint(combine), 0, 255).astype(np.uint8)
