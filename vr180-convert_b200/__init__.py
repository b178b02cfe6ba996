"""vr180_convert_b200 -- B200-native reprojection hot path of vr180-convert behind the reference's Python API.

Same public names as /root/reference/src/vr180_convert/__init__.py:2-31, plus the batched / device-resident
entry points (`SbsWarper`, `shard_range`) used for video and multi-GPU runs.
"""
__version__ = "0.1.0"

from .quat import from_euler_angles, from_rotation_vector, quaternion, rotate_vectors
from .remapper import apply, apply_lr, get_map, get_radius_smart, lr_frame, lr_frames, match_lr, pinned_empty, remap_maps, set_codec, set_device
from .transformer import (
    DenormalizeTransformer,
    EquirectangularDecoder,
    EquirectangularEncoder,
    Euclidean3DRotator,
    Euclidean3DTransformer,
    FisheyeDecoder,
    FisheyeEncoder,
    InverseTransformer,
    MultiTransformer,
    NormalizeTransformer,
    PolarRollTransformer,
    PolynomialScaler,
    RectilinearDecoder,
    TransformerBase,
    ZoomTransformer,
    equidistant_from_3d,
    equidistant_to_3d,
    get_radius,
)
from .shard import gather_shards, max_over_ranks, shard_range
from .video import SbsWarper

__all__ = [
    "TransformerBase", "ZoomTransformer", "MultiTransformer", "NormalizeTransformer", "PolarRollTransformer",
    "DenormalizeTransformer", "FisheyeDecoder", "FisheyeEncoder", "EquirectangularEncoder", "EquirectangularDecoder",
    "Euclidean3DRotator", "Euclidean3DTransformer", "InverseTransformer", "PolynomialScaler", "RectilinearDecoder",
    "apply", "apply_lr", "get_map", "get_radius", "get_radius_smart", "lr_frame", "lr_frames", "match_lr", "pinned_empty", "remap_maps", "set_codec", "set_device",
    "equidistant_to_3d", "equidistant_from_3d", "quaternion", "from_rotation_vector", "from_euler_angles",
    "rotate_vectors", "SbsWarper", "shard_range", "max_over_ranks", "gather_shards",
]
