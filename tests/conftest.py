"""pytest configuration: `gpu` marker, repo root on sys.path, golden-fixture loaders."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden_maps():
    z = np.load(GOLDEN / "maps.npz")
    meta = json.loads(str(z["meta"]))
    return z, meta


@pytest.fixture(scope="session")
def golden_remap():
    return np.load(GOLDEN / "remap.npz")


@pytest.fixture(scope="session")
def golden_apply():
    return np.load(GOLDEN / "apply.npz")


@pytest.fixture(scope="session")
def golden_radius():
    return json.loads((GOLDEN / "radius.json").read_text())


@pytest.fixture(scope="session")
def golden_fullsize():
    return np.load(GOLDEN / "maps_fullsize_samples.npz")


@pytest.fixture(scope="session")
def golden_cfg1():
    """BASELINE.json configs[0]: the reference's `lr test.jpg test.jpg` run (tests/golden/make_golden.py:make_cfg1)."""
    z = np.load(GOLDEN / "cfg1.npz")
    return z, json.loads(str(z["meta"])), GOLDEN / "cfg1_test.jpg"


@pytest.fixture(scope="session")
def golden_merge_match():
    """apply_lr(merge=True) outputs and match_lr vectors of the unmodified reference (make_golden.py merge)."""
    return np.load(GOLDEN / "merge_match.npz")


def disc_frame(h: int, w: int, seed: int, margin: int = 8) -> np.ndarray:
    """Synthetic fisheye frame of SURVEY.md §8(d): uniform random bytes inside the disc, zeros outside."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    yy, xx = np.ogrid[:h, :w]
    r = min(h, w) // 2 - margin
    img[(xx - w // 2) ** 2 + (yy - h // 2) ** 2 > r * r] = 0
    return img
