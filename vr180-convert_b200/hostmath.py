"""NumPy evaluation of lowered chain ops on ARBITRARY coordinate arrays (host side).

This backs the public array API `TransformerBase.transform(x, y)` / `inverse_transform(x, y)` (reference:
transformer.py:24-69), which callers use on short point lists (e.g. calibration points, remapper.py:291-320).
It is not on the image path: `get_map`, `apply` and `apply_lr` evaluate lowerable chains with the CUDA kernels
and never call into this module for them.
"""
from __future__ import annotations

import numpy as np

HALF_PI = np.pi / 2


def _polar(x, y):
    r = np.sqrt(x * x + y * y)
    roll = np.arctan2(y, x)
    return r, np.cos(roll), np.sin(roll)


def _to_vec3(x, y):
    phi = np.arctan2(x, y)
    th = np.sqrt(x * x + y * y)
    st = np.sin(th)
    return st * np.sin(phi), st * np.cos(phi), np.cos(th)


def _from_vec3(vx, vy, vz):
    th = np.arccos(vz)
    phi = np.arctan2(vx, vy)
    return th * np.sin(phi), th * np.cos(phi)


_R_TO_THETA = {
    "rectilinear": np.arctan,
    "stereographic": lambda r: 2 * np.arctan(r),
    "equidistant": lambda r: r * HALF_PI,
    "equisolid": lambda r: 2 * np.arcsin(r / np.sqrt(2)),
    "orthographic": np.arcsin,
}
_THETA_TO_R = {
    "rectilinear": np.tan,
    "stereographic": lambda t: 2 * np.tan(t / 2),
    "equidistant": lambda t: t / HALF_PI,
    "equisolid": lambda t: np.sqrt(2) * np.sin(t / 2),
    "orthographic": np.sin,
}


def run_ops(ops, x, y):
    x = np.asarray(x)
    y = np.asarray(y)
    for op in ops:
        kind = op[0]
        if kind == "normalize":
            (cx, cy), s = op[1], op[2]
            x, y = (x - cx) / s * 2, (y - cy) / s * 2
        elif kind == "denormalize":
            (sx, sy), (cx, cy) = op[1], op[2]
            x, y = x * sx + cx, y * sy + cy
        elif kind == "denormalize_inv":
            (sx, sy), (cx, cy) = op[1], op[2]
            x, y = (x - cx) / sx, (y - cy) / sy
        elif kind == "zoom":
            x, y = x / op[1], y / op[1]
        elif kind == "zoom_inv":
            x, y = x * op[1], y * op[1]
        elif kind == "equirect_enc":
            lat, lon = ((y, x) if op[1] else (x, y))
            lat, lon = lat * HALF_PI, lon * HALF_PI
            a, b, c = np.cos(lat) * np.sin(lon), np.sin(lat), np.cos(lat) * np.cos(lon)
            x, y = _from_vec3(a, b, c) if op[1] else _from_vec3(b, a, c)
        elif kind == "equirect_dec":
            vx, vy, vz = _to_vec3(x, y)
            if op[1]:
                x, y = np.arctan2(vx, vz) / HALF_PI, np.arcsin(vy) / HALF_PI
            else:
                x, y = np.arcsin(vx) / HALF_PI, np.arctan2(vy, vz) / HALF_PI
        elif kind in ("fisheye_enc", "fisheye_dec", "rectilinear_dec", "rectilinear_dec_inv", "poly"):
            r, cr, sr = _polar(x, y)
            if kind == "fisheye_enc":
                r = _R_TO_THETA[op[1]](r)
            elif kind == "fisheye_dec":
                r = _THETA_TO_R[op[1]](r)
            elif kind == "rectilinear_dec":
                r = np.tan(r) * op[1]
            elif kind == "rectilinear_dec_inv":
                r = np.arctan(r / op[1])
            else:
                r = np.polyval(np.asarray(op[1], dtype=np.float64)[::-1], r)
            x, y = r * cr, r * sr
        elif kind == "rot3":
            m = np.asarray(op[1], dtype=np.float64).reshape(3, 3)
            vx, vy, vz = _to_vec3(x, y)
            x, y = _from_vec3(m[0, 0] * vx + m[0, 1] * vy + m[0, 2] * vz, m[1, 0] * vx + m[1, 1] * vy + m[1, 2] * vz,
                              m[2, 0] * vx + m[2, 1] * vy + m[2, 2] * vz)
        else:
            raise ValueError(f"unknown chain op {kind!r}")
    return x, y
