"""Helpers shared by the tests."""
import hashlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
FIX = ROOT / "oracle" / "_ref" / "fixtures"


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def png(name: str) -> np.ndarray:
    from PIL import Image
    return np.ascontiguousarray(np.asarray(Image.open(GOLD / name)))


def fixture_raw(name: str, shape):
    p = FIX / f"{name}.raw"
    if not p.exists():
        return None
    return np.fromfile(p, np.uint8).reshape(shape)


SYNTH_SHAPES = [(1, 1, 1), (2, 7, 2), (3, 3, 3), (3, 64, 4), (17, 3, 5), (64, 64, 6), (257, 333, 7),
                (480, 2037, 8), (1080, 1920, 9), (2160, 3840, 10)]   # (h, w, seed) as in make_golden.py
