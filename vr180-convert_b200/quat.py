"""Minimal quaternion value type for `Euclidean3DRotator(rotation=...)`.

The reference takes numpy-quaternion objects (transformer.py:10, 672).  numpy-quaternion is an optional
third-party C extension that is not required here: anything exposing `.w .x .y .z` (numpy-quaternion objects
included), or a 4-sequence (w, x, y, z), is accepted by `as_wxyz`.  Scalar-first convention; a rotation acts as
v' = R(q) v with R normalised by |q|^2, so non-unit quaternions (cli.py:308-311 builds them) are pure rotations.
"""
from __future__ import annotations

import math
from typing import Any, Sequence

import numpy as np


class quaternion:  # noqa: N801 - mirrors the numpy-quaternion type name used in transformer expressions
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, w: float = 1.0, x: float = 0.0, y: float = 0.0, z: float = 0.0):
        self.w, self.x, self.y, self.z = float(w), float(x), float(y), float(z)

    @property
    def components(self) -> np.ndarray:
        return np.array([self.w, self.x, self.y, self.z], dtype=np.float64)

    def norm(self) -> float:
        return float(np.dot(self.components, self.components))

    def conj(self) -> "quaternion":
        return quaternion(self.w, -self.x, -self.y, -self.z)

    conjugate = conj

    def inverse(self) -> "quaternion":
        n = self.norm()
        return quaternion(self.w / n, -self.x / n, -self.y / n, -self.z / n)

    def __neg__(self) -> "quaternion":
        return quaternion(-self.w, -self.x, -self.y, -self.z)

    def __add__(self, other: Any) -> "quaternion":
        if hasattr(other, "w"):
            return quaternion(self.w + other.w, self.x + other.x, self.y + other.y, self.z + other.z)
        return quaternion(self.w + float(other), self.x, self.y, self.z)

    __radd__ = __add__

    def __mul__(self, other: Any) -> "quaternion":
        if hasattr(other, "w"):
            aw, ax, ay, az = self.w, self.x, self.y, self.z
            bw, bx, by, bz = other.w, other.x, other.y, other.z
            return quaternion(aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                              aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw)
        k = float(other)
        return quaternion(self.w * k, self.x * k, self.y * k, self.z * k)

    def __rmul__(self, other: Any) -> "quaternion":
        return self * float(other)

    def __repr__(self) -> str:
        return f"quaternion({self.w!r}, {self.x!r}, {self.y!r}, {self.z!r})"

    def __eq__(self, other: Any) -> bool:
        return hasattr(other, "w") and tuple(as_wxyz(other)) == tuple(as_wxyz(self))

    def __hash__(self) -> int:
        return hash((self.w, self.x, self.y, self.z))


def as_wxyz(q: Any) -> tuple[float, float, float, float]:
    if hasattr(q, "w") and hasattr(q, "z"):
        return float(q.w), float(q.x), float(q.y), float(q.z)
    a = np.asarray(q, dtype=np.float64).reshape(-1)
    if a.size != 4:
        raise TypeError("rotation must be a quaternion (w, x, y, z)")
    return float(a[0]), float(a[1]), float(a[2]), float(a[3])


def rotation_matrix(q: Any) -> np.ndarray:
    """3x3 matrix of v -> q v q^-1 (what numpy-quaternion's rotate_vectors multiplies by)."""
    w, x, y, z = as_wxyz(q)
    n = w * w + x * x + y * y + z * z
    if n == 0.0:
        raise ZeroDivisionError("zero quaternion has no rotation")
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
                     [s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w)],
                     [s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)]], dtype=np.float64)


def rotate_vectors(q: Any, v: np.ndarray, axis: int = -1) -> np.ndarray:
    v = np.asarray(v, dtype=np.float64)
    return np.moveaxis(np.tensordot(rotation_matrix(q), v, axes=(-1, axis)), 0, axis)


def from_rotation_vector(rot: Sequence[float]) -> quaternion:
    r = np.asarray(rot, dtype=np.float64)
    angle = float(np.linalg.norm(r))
    if angle == 0.0:
        return quaternion(1.0, 0.0, 0.0, 0.0)
    k = math.sin(angle / 2) / angle
    return quaternion(math.cos(angle / 2), r[0] * k, r[1] * k, r[2] * k)


def from_euler_angles(alpha: float, beta: float, gamma: float) -> quaternion:
    """ZYZ convention: exp(alpha z/2) exp(beta y/2) exp(gamma z/2)."""
    return quaternion(math.cos(beta / 2) * math.cos((alpha + gamma) / 2), -math.sin(beta / 2) * math.sin((alpha - gamma) / 2),
                      math.sin(beta / 2) * math.cos((alpha - gamma) / 2), math.cos(beta / 2) * math.sin((alpha + gamma) / 2))
