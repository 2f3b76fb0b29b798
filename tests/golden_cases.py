"""Replay helpers for tests/golden/*.bin (made by oracle/make_golden.py from the unmodified reference)."""
import os

import numpy as np

from refutil import GOLDEN_DIR, read_bundle


def load(name):
    b = read_bundle(os.path.join(GOLDEN_DIR, name))
    out = {}
    for k, (t, ne, data) in b.items():
        out[k] = (t, ne, np.frombuffer(data, dtype=np.uint8).copy())
    return out


def f32(entry):
    t, ne, raw = entry
    shape = [d for d in reversed(ne)]
    while len(shape) > 1 and shape[0] == 1:
        shape = shape[1:]
    return raw.view(np.float32).reshape(shape)


def raw2d(entry, rows):
    return entry[2].reshape(rows, -1)
