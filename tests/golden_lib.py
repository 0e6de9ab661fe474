"""Loader for tests/golden/*.npz (made by tests/golden/make_golden.py from the reference's own lines)."""
import glob
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind):
    return sorted(os.path.basename(p)[len(kind) + 1:-4] for p in glob.glob(os.path.join(GOLDEN, f"{kind}_*.npz")))


def load(kind, name):
    return np.load(os.path.join(GOLDEN, f"{kind}_{name}.npz"))


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)
