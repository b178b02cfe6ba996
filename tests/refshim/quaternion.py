"""TEST-ONLY stand-in for numpy-quaternion (pinned 2023.0.4 in the reference's poetry.lock:947-948).

numpy-quaternion is not installed in this image and there is no network, but the reference imports it at
module import time (/root/reference/src/vr180_convert/transformer.py:10, remapper.py:10). This shim supplies
exactly the names the reference and its tests use so that the UNMODIFIED reference can be imported to
generate golden vectors (tests/golden/make_golden.py) and to validate the oracle.  It is never imported by
the product package.

Semantics restated from numpy-quaternion's documentation (scalar-first w,x,y,z; rotate_vectors(q, v) = R(q) v
with R normalised by |q|^2) and cross-checked against scipy.spatial.transform.Rotation in
tests/test_oracle_golden.py::test_quaternion_shim_matches_scipy.
"""
from __future__ import annotations

import numpy as np


class quaternion:  # noqa: N801  (name fixed by the library being stood in for)
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, w=0.0, x=0.0, y=0.0, z=0.0):
        self.w, self.x, self.y, self.z = float(w), float(x), float(y), float(z)

    # --- algebra -----------------------------------------------------------------------------------------
    def norm(self) -> float:  # numpy-quaternion's .norm() is the squared magnitude ("Cayley norm")
        return self.w * self.w + self.x * self.x + self.y * self.y + self.z * self.z

    def abs(self) -> float:
        return float(np.sqrt(self.norm()))

    def conj(self) -> "quaternion":
        return quaternion(self.w, -self.x, -self.y, -self.z)

    conjugate = conj

    def inverse(self) -> "quaternion":
        n = self.norm()
        return quaternion(self.w / n, -self.x / n, -self.y / n, -self.z / n)

    def __neg__(self):
        return quaternion(-self.w, -self.x, -self.y, -self.z)

    def __add__(self, o):
        if isinstance(o, quaternion):
            return quaternion(self.w + o.w, self.x + o.x, self.y + o.y, self.z + o.z)
        return quaternion(self.w + float(o), self.x, self.y, self.z)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        if isinstance(o, quaternion):
            a, b = self, o
            return quaternion(
                a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
                a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
                a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
            )
        s = float(o)
        return quaternion(self.w * s, self.x * s, self.y * s, self.z * s)

    def __rmul__(self, o):
        s = float(o)
        return quaternion(self.w * s, self.x * s, self.y * s, self.z * s)

    def __truediv__(self, o):
        if isinstance(o, quaternion):
            return self * o.inverse()
        return self * (1.0 / float(o))

    @property
    def components(self):
        return np.array([self.w, self.x, self.y, self.z])

    def __repr__(self):
        return f"quaternion({self.w!r}, {self.x!r}, {self.y!r}, {self.z!r})"

    def __eq__(self, o):
        return isinstance(o, quaternion) and bool(np.all(self.components == o.components))

    def __hash__(self):
        return hash(tuple(self.components))


def as_float_array(q):
    return q.components


def as_quat_array(a):
    a = np.asarray(a, dtype=float)
    return quaternion(*a[..., :4].reshape(-1)[:4])


def as_rotation_matrix(q):
    w, x, y, z = q.w, q.x, q.y, q.z
    n = q.norm()
    return np.array(
        [
            [1 - 2 * (y * y + z * z) / n, 2 * (x * y - z * w) / n, 2 * (x * z + y * w) / n],
            [2 * (x * y + z * w) / n, 1 - 2 * (x * x + z * z) / n, 2 * (y * z - x * w) / n],
            [2 * (x * z - y * w) / n, 2 * (y * z + x * w) / n, 1 - 2 * (x * x + y * y) / n],
        ]
    )


def rotate_vectors(q, v, axis=-1):
    v = np.asarray(v, dtype=float)
    m = as_rotation_matrix(q)
    return np.moveaxis(np.tensordot(m, v, axes=(-1, axis)), 0, axis)


def from_rotation_vector(r):
    r = np.asarray(r, dtype=float)
    a = float(np.linalg.norm(r))
    if a == 0.0:
        return quaternion(1, 0, 0, 0)
    s = np.sin(a / 2) / a
    return quaternion(np.cos(a / 2), *(r * s))


def from_euler_angles(alpha, beta=None, gamma=None):
    if beta is None:
        alpha, beta, gamma = alpha
    return quaternion(
        np.cos(beta / 2) * np.cos((alpha + gamma) / 2),
        -np.sin(beta / 2) * np.sin((alpha - gamma) / 2),
        np.sin(beta / 2) * np.cos((alpha - gamma) / 2),
        np.cos(beta / 2) * np.sin((alpha + gamma) / 2),
    )


def allclose(a, b, rtol=4 * np.finfo(float).eps, atol=0.0, equal_nan=False, verbose=False):
    return bool(np.allclose(a.components, b.components, rtol=rtol, atol=atol, equal_nan=equal_nan))


one = quaternion(1, 0, 0, 0)
x = quaternion(0, 1, 0, 0)
y = quaternion(0, 0, 1, 0)
z = quaternion(0, 0, 0, 1)
