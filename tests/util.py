"""Shared helpers for tests: golden vectors, conversions."""
import json
import os

import numpy as np

import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        return json.load(f)


def ints(hexes):
    return [int(h, 16) for h in hexes]


def pt_from_golden(p):
    return None if p is None else (int(p[0], 16), int(p[1], 16))


def frs(vs):
    return np.frombuffer(b"".join(int(v % P.R_MOD).to_bytes(32, "little") for v in vs), dtype=np.uint64).reshape(-1, 4).copy()


def fr1(v):
    return frs([v])[0]


def to_ints(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def g1s(pts):
    return np.frombuffer(b"".join(P.g1_to_bytes(p) for p in pts), dtype=np.uint64).reshape(-1, 12).copy()


def g1_tuple(a):
    return P.g1_from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
