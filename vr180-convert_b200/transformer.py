"""Transformer algebra with the reference's public names, lowered to a chain descriptor for the CUDA kernels.

Mirrors /root/reference/src/vr180_convert/transformer.py (class names, constructor fields, `a * b` composition,
`transform` / `inverse_transform(x, y, **kw) -> (x, y)` on NumPy arrays, error types), but the classes here are
thin *descriptions*: each recognised transformer knows how to `lower()` itself to op tuples
(include/vr180_b200.h `vr180_op_code`), and the per-pixel arithmetic for images runs in csrc/chain.cuh.
User-defined subclasses (README.md:204-219) have no lowering; chains containing one are evaluated once on the
host by the user's own NumPy code and enter the GPU through the LUT path (remapper.get_map).
"""
from __future__ import annotations

import warnings
from abc import ABCMeta, abstractmethod
from typing import Any, Generic, Literal, Sequence, TypeVar

import attrs
import numpy as np
from numpy.typing import NDArray
from sklearn.base import BaseEstimator, TransformerMixin

from . import hostmath
from .quat import quaternion, rotation_matrix

Ops = "list[tuple]"
_MAPPINGS = ("rectilinear", "stereographic", "equidistant", "equisolid", "orthographic")


def _unknown_mapping(name: Any) -> ValueError:
    return ValueError(f"Unknown mapping type: {name}, should be one of "
                      "'rectilinear', 'stereographic', 'equidistant', 'equisolid', 'orthographic'.")


class TransformerBase(BaseEstimator, TransformerMixin, metaclass=ABCMeta):
    """Coordinate transformer: new_image[(x, y)] = old_image[transform(x, y)]  (transformer.py:14-81).
    As in the reference the base is a scikit-learn estimator (transformer.py:11, :14-18): `get_params` /
    `set_params` / `clone` work on every transformer, and `repr` of the attrs subclasses is the cache key of
    a lowered chain (remapper.lowered_chain)."""

    @abstractmethod
    def transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        ...

    @abstractmethod
    def inverse_transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        ...

    def lower(self, shape: tuple[int, int] | None = None, inverse: bool = False) -> "list[tuple] | None":
        """Op tuples for the CUDA chain evaluator, or None when this transformer is opaque Python.
        `shape` = (rows, cols) of the coordinate grid (only NormalizeTransformer needs it)."""
        return None

    def __mul__(self, other: "TransformerBase") -> "MultiTransformer":
        left = self.transformers if isinstance(self, MultiTransformer) else [self]
        right = other.transformers if isinstance(other, MultiTransformer) else [other]
        return MultiTransformer(transformers=[*left, *right])


T = TypeVar("T", bound=TransformerBase)


def _pristine(obj: Any, cls: type, *methods: str) -> bool:
    """True when `obj` still uses `cls`'s own implementations (a subclass may have overridden the math)."""
    return all(getattr(type(obj), m) is getattr(cls, m) for m in methods)


class _Lowered(TransformerBase):
    """Shared array API for transformers that have a lowering: evaluate the ops with NumPy (points API)."""

    def _ops(self, x: NDArray, inverse: bool) -> list[tuple]:
        shape = np.shape(x)
        ops = self.lower(shape=(shape[0], shape[1]) if len(shape) >= 2 else None, inverse=inverse)
        if ops is None:
            raise NotImplementedError(f"{type(self).__name__} cannot be evaluated")
        return ops

    def transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        return hostmath.run_ops(self._ops(x, False), x, y)

    def inverse_transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        return hostmath.run_ops(self._ops(x, True), x, y)


@attrs.define()
class MultiTransformer(TransformerBase):
    """Sequential composition (transformer.py:87-105): forward in list order, inverse in reverse order."""

    transformers: list[TransformerBase]

    def transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        for t in self.transformers:
            x, y = t.transform(x, y, **kwargs)
        return x, y

    def inverse_transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        for t in self.transformers[::-1]:
            x, y = t.inverse_transform(x, y, **kwargs)
        return x, y

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, MultiTransformer, "transform", "inverse_transform"):
            return None  # a subclass re-defined the composition in Python: opaque, host LUT route
        out: list[tuple] = []
        for t in (self.transformers[::-1] if inverse else self.transformers):
            ops = t.lower(shape=shape, inverse=inverse)
            if ops is None:
                return None
            out.extend(ops)
        return out


@attrs.define()
class NormalizeTransformer(_Lowered):
    """Pixel grid -> [-1, 1] (transformer.py:143-177).  centre defaults to (cols/2, rows/2), scale to min(cols, rows)."""

    center: tuple[float, float] | None = None
    scale: tuple[float, float] | Literal["min", "max"] | None = None

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, NormalizeTransformer, "transform", "inverse_transform"):
            return None
        if shape is None:
            raise ValueError("NormalizeTransformer needs a 2-D coordinate grid")
        rows, cols = shape
        center = self.center or (cols / 2, rows / 2)
        if self.scale in ("min", None):
            scale: Any = min(cols, rows)
        elif self.scale == "max":
            scale = max(cols, rows)
        else:
            scale = self.scale
        if inverse:  # transformer.py:175-176 indexes scale[0] / scale[1]; a scalar raises TypeError there too
            return [("denormalize", (scale[0], scale[1]), (center[0], center[1]))]
        if isinstance(scale, (tuple, list)):
            raise TypeError("a (sx, sy) scale is only usable by inverse_transform, as in the reference")
        return [("normalize", (float(center[0]), float(center[1])), float(scale))]


@attrs.define()
class DenormalizeTransformer(_Lowered):
    """[-1, 1] -> source pixels (transformer.py:188-213)."""

    scale: tuple[float, float]
    center: tuple[float, float]

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, DenormalizeTransformer, "transform", "inverse_transform"):
            return None
        kind = "denormalize_inv" if inverse else "denormalize"
        return [(kind, (float(self.scale[0]), float(self.scale[1])), (float(self.center[0]), float(self.center[1])))]


@attrs.define()
class PolarRollTransformer(TransformerBase):
    """Radial transformer: (x, y) -> (theta, roll) -> user function -> (x, y)  (transformer.py:216-286).
    Subclasses implementing `transform_polar` in Python are opaque to the GPU (LUT path)."""

    @abstractmethod
    def transform_polar(self, theta: NDArray, roll: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        ...

    def inverse_transform_polar(self, theta: NDArray, roll: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        raise NotImplementedError(f"{type(self).__name__} does not support inverse transform.")

    def _through_polar(self, fn, x, y, kwargs):
        theta = np.sqrt(x**2 + y**2)
        roll = np.arctan2(y, x)
        theta, roll = fn(theta, roll, **kwargs)
        return theta * np.cos(roll), theta * np.sin(roll)

    def transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        return self._through_polar(self.transform_polar, x, y, kwargs)

    def inverse_transform(self, x: NDArray, y: NDArray, **kwargs: Any) -> tuple[NDArray, NDArray]:
        return self._through_polar(self.inverse_transform_polar, x, y, kwargs)


_SENSOR_WIDTHS_MM = {
    "35mm": 36.0, "APS-H": 27.90, "APS-C": 23.6, "APS-C-Canon": 22.30, "MFT": 17.30, "1": 13.20, "1/1.12": 11.43,
    "1/1.2": 10.67, "1/1.33": 9.6, "1/1.6": 8.08, "1/1.7": 7.60, "1/1.8": 7.18, "1/2": 6.40, "1/2.3": 6.17,
}


@attrs.define()
class RectilinearDecoder(PolarRollTransformer):
    """Ordinary (rectilinear) lens with a focal length in mm (transformer.py:289-347)."""

    focal_length: float
    sensor_width: Literal["35mm", "APS-H", "APS-C", "APS-C-Canon", "Foveon", "MFT"] | str | float = "35mm"

    @property
    def sensor_width_mm(self) -> float:
        if self.sensor_width in ("35mm", "APS-C", "1/2.3"):
            warnings.warn(
                "Sensor size may vary by about 0.2 mm depending on the camera model. "
                "To get very accurate results, consider setting the sensor width in mm manually.",
                UserWarning, stacklevel=2)
        if isinstance(self.sensor_width, str):
            return _SENSOR_WIDTHS_MM[self.sensor_width]
        return self.sensor_width

    @property
    def factor(self) -> float:
        return 2 * self.focal_length / self.sensor_width_mm

    def transform_polar(self, theta, roll, **kwargs):
        return np.tan(theta) * self.factor, roll

    def inverse_transform_polar(self, theta, roll, **kwargs):
        return np.arctan(theta / self.factor), roll

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, RectilinearDecoder, "transform_polar", "inverse_transform_polar", "transform",
                         "inverse_transform"):
            return None
        return [("rectilinear_dec_inv" if inverse else "rectilinear_dec", float(self.factor))]


@attrs.define()
class FisheyeEncoder(PolarRollTransformer):
    """Fisheye radius [-1, 1] -> angle (transformer.py:350-397); five classical mapping functions."""

    mapping_type: Literal["rectilinear", "stereographic", "equidistant", "equisolid", "orthographic"]

    def _check(self) -> str:
        if self.mapping_type not in _MAPPINGS:
            raise _unknown_mapping(self.mapping_type)
        return self.mapping_type

    def transform_polar(self, theta, roll, **kwargs):
        return hostmath._R_TO_THETA[self._check()](theta), roll

    def inverse_transform_polar(self, theta, roll, **kwargs):
        return hostmath._THETA_TO_R[self._check()](theta), roll

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, FisheyeEncoder, "transform_polar", "inverse_transform_polar", "transform",
                         "inverse_transform"):
            return None
        return [("fisheye_dec" if inverse else "fisheye_enc", self._check())]


@attrs.define()
class InverseTransformer(TransformerBase, Generic[T]):
    """Swaps transform and inverse_transform of the wrapped transformer (transformer.py:400-415)."""

    transformer: T

    def transform(self, x, y, **kwargs):
        return self.transformer.inverse_transform(x, y, **kwargs)

    def inverse_transform(self, x, y, **kwargs):
        return self.transformer.transform(x, y, **kwargs)

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, InverseTransformer, "transform", "inverse_transform"):
            return None
        return self.transformer.lower(shape=shape, inverse=not inverse)


def FisheyeDecoder(  # noqa: N802 - factory named like a class, as in the reference (transformer.py:418-437)
    mapping_type: Literal["rectilinear", "stereographic", "equidistant", "equisolid", "orthographic"],
) -> InverseTransformer[FisheyeEncoder]:
    return InverseTransformer(FisheyeEncoder(mapping_type))


def _default_coefs() -> list[float]:
    return [0, 1]


@attrs.define()
class PolynomialScaler(PolarRollTransformer):
    """theta' = c0 + c1 theta + c2 theta^2 + ... (transformer.py:440-458); forward only."""

    coefs_reverse: Sequence[float] = attrs.field(factory=_default_coefs)

    def transform_polar(self, theta, roll, **kwargs):
        return np.polyval(np.asarray(self.coefs_reverse, dtype=np.float64)[::-1], theta), roll

    def inverse_transform_polar(self, theta, roll, **kwargs):
        raise NotImplementedError("PolynomialScaler does not support inverse transform.")

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, PolynomialScaler, "transform_polar", "transform"):
            return None
        if inverse:
            raise NotImplementedError("PolynomialScaler does not support inverse transform.")
        if len(self.coefs_reverse) > 12:
            return None  # kernel holds up to 12 coefficients; larger polynomials take the host/LUT route
        return [("poly", [float(c) for c in self.coefs_reverse])]


@attrs.define()
class ZoomTransformer(_Lowered):
    """x / scale (transformer.py:461-480)."""

    scale: float

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, ZoomTransformer, "transform", "inverse_transform"):
            return None
        return [("zoom_inv" if inverse else "zoom", float(self.scale))]


def equidistant_to_3d(x: NDArray, y: NDArray) -> NDArray:
    """Equidistant angle coordinates -> unit vectors, z forward (transformer.py:483-508)."""
    vx, vy, vz = hostmath._to_vec3(np.asarray(x), np.asarray(y))
    return np.stack([vx, vy, vz], axis=-1)


def equidistant_from_3d(v: NDArray) -> tuple[NDArray, NDArray]:
    """Unit vectors -> equidistant angle coordinates (transformer.py:511-530)."""
    v = np.asarray(v)
    return hostmath._from_vec3(v[..., 0], v[..., 1], v[..., 2])


@attrs.define()
class EquirectangularEncoder(_Lowered):
    """Equirectangular [-1, 1]^2 (lat/lon * 2/pi) -> equidistant angle coordinates (transformer.py:533-584)."""

    is_latitude_y: bool = True

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, EquirectangularEncoder, "transform", "inverse_transform"):
            return None
        return [("equirect_dec" if inverse else "equirect_enc", bool(self.is_latitude_y))]


def EquirectangularDecoder(is_latitude_y: bool = True) -> InverseTransformer[EquirectangularEncoder]:  # noqa: N802
    return InverseTransformer(EquirectangularEncoder(is_latitude_y))


@attrs.define()
class Euclidean3DTransformer(TransformerBase):
    """Acts on the unit vector of each coordinate (transformer.py:607-665).  As in the reference,
    inverse_transform applies `transform_v` too (transformer.py:659-665)."""

    @abstractmethod
    def transform_v(self, v: NDArray) -> NDArray:
        ...

    @abstractmethod
    def inverse_transform_v(self, v: NDArray) -> NDArray:
        ...

    def transform(self, x, y, **kwargs):
        return equidistant_from_3d(self.transform_v(equidistant_to_3d(x, y)))

    def inverse_transform(self, x, y, **kwargs):
        return equidistant_from_3d(self.transform_v(equidistant_to_3d(x, y)))


@attrs.define()
class Euclidean3DRotator(Euclidean3DTransformer):
    """Rotation by a quaternion (transformer.py:668-679).  Accepts numpy-quaternion objects, `quat.quaternion`,
    or (w, x, y, z)."""

    rotation: Any

    def transform_v(self, v):
        return np.moveaxis(np.tensordot(rotation_matrix(self.rotation), np.asarray(v), axes=(-1, -1)), 0, -1)

    def inverse_transform_v(self, v):
        return np.moveaxis(np.tensordot(rotation_matrix(self.rotation).T, np.asarray(v), axes=(-1, -1)), 0, -1)

    def lower(self, shape=None, inverse=False):
        if not _pristine(self, Euclidean3DRotator, "transform_v", "transform", "inverse_transform"):
            return None
        return [("rot3", rotation_matrix(self.rotation).reshape(-1).tolist())]  # same matrix both ways, see class doc


# ---------------------------------------------------------------------------------------------------------
# get_radius (transformer.py:108-140) -- device reduction
# ---------------------------------------------------------------------------------------------------------
def get_radius(input: NDArray, *, threshold: int = 10) -> float:  # noqa: A002 - reference argument name
    """Fisheye-circle radius from the black / non-black transitions of the centre row (cols > rows) or centre
    column.  Runs csrc/kernels.cu:k_get_radius; raises IndexError when a transition is missing, like the
    reference's `np.where(...)[0][0]`."""
    from .remapper import _device_radius  # late import: remapper imports this module

    radii = _device_radius([input], threshold)
    return radii[0]


__all__ = [
    "TransformerBase", "MultiTransformer", "NormalizeTransformer", "DenormalizeTransformer", "PolarRollTransformer",
    "RectilinearDecoder", "FisheyeEncoder", "FisheyeDecoder", "InverseTransformer", "PolynomialScaler",
    "ZoomTransformer", "EquirectangularEncoder", "EquirectangularDecoder", "Euclidean3DTransformer",
    "Euclidean3DRotator", "equidistant_to_3d", "equidistant_from_3d", "get_radius", "quaternion",
]
