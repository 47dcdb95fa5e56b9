"""CPU check of the device arithmetic *text* (ff.cuh / g1.cuh) through its host emulation of the
PTX carry chains, against the big-integer oracle.  The GPU tests repeat this on the real PTX path."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import pyref as P

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def L():
    d = os.path.join(HERE, "host_arith")
    so = os.path.join(d, "libhost_arith.so")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(d, "host_arith.cpp")])
    return ctypes.CDLL(so)


def u32(v, n):
    return np.frombuffer(int(v).to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def toint(a):
    return int.from_bytes(a.tobytes(), "little")


def pp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_field_ops(L, name):
    mod, n, fn = (P.R_MOD, 8, L.h_fr_op) if name == "fr" else (P.Q_MOD, 12, L.h_fq_op)
    rng = random.Random(1)
    edge = [0, 1, 2, mod - 1, mod - 2, (1 << (32 * n)) % mod, (1 << 32) - 1, 1 << 32, mod >> 1]
    vals = edge + [rng.randrange(mod) for _ in range(400)]
    rinv = pow(1 << (32 * n), -1, mod)
    pairs = [(a, b) for a in edge for b in edge] + [(vals[i], vals[(i * 7 + 3) % len(vals)]) for i in range(len(vals))]
    for a, b in pairs:
        for op, exp in ((0, (a + b) % mod), (1, (a - b) % mod), (2, a * b % mod), (4, a * b * rinv % mod)):
            o = np.zeros(n, dtype=np.uint32)
            fn(op, pp(u32(a, n)), pp(u32(b, n)), pp(o))
            assert toint(o) == exp, (name, op, hex(a), hex(b))
    for a in vals[:24]:
        o = np.zeros(n, dtype=np.uint32)
        fn(3, pp(u32(a, n)), pp(u32(a, n)), pp(o))
        assert toint(o) == pow(a, mod - 2, mod)


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_dedicated_squaring(L, name):
    """Fp::sqr (N(N+1)/2 products + interleaved reduction) against big integers: canonical inputs and raw Montgomery
    inputs up to the container limit used inside the point formulas."""
    mod, n, fn = (P.R_MOD, 8, L.h_fr_op) if name == "fr" else (P.Q_MOD, 12, L.h_fq_op)
    rng = random.Random(7)
    R = 1 << (32 * n)
    rinv = pow(R, -1, mod)
    top = (1 << 31) - 1
    edge = [0, 1, 2, 3, mod - 1, mod - 2, R % mod, (1 << 32) - 1, 1 << 32, (1 << 31), mod >> 1, (mod >> 1) + 1,
            int("ffffffff" * (n - 1), 16), int("80000000" * (n - 1), 16) % mod, int("7fffffff" + "ffffffff" * (n - 1), 16) % mod]
    vals = edge + [rng.randrange(mod) for _ in range(600)] + [(rng.randrange(1 << 31) << (32 * (n - 1))) % mod for _ in range(50)]
    for a in vals:
        o = np.zeros(n, dtype=np.uint32)
        fn(5, pp(u32(a, n)), pp(u32(a, n)), pp(o))
        assert toint(o) == a * a % mod, (name, hex(a))
        o2 = np.zeros(n, dtype=np.uint32)
        fn(6, pp(u32(a, n)), pp(u32(a, n)), pp(o2))  # raw Montgomery square: a*a/R mod p
        assert toint(o2) == a * a * rinv % mod, (name, "raw", hex(a))
        o3 = np.zeros(n, dtype=np.uint32)
        fn(4, pp(u32(a, n)), pp(u32(a, n)), pp(o3))
        assert np.array_equal(o2, o3)


def g1b(pt):
    return np.frombuffer(P.g1_to_bytes(pt), dtype=np.uint32).copy()


def g1t(a):
    return P.g1_from_bytes(a.tobytes())


def test_g1_ops(L):
    rng = random.Random(2)
    G = P.G1_GEN
    pts = [P.g1_mul(G, rng.randrange(P.R_MOD)) for _ in range(12)]
    cases = [
        pts,
        [pts[0], pts[0]],
        [pts[0], P.g1_neg(pts[0]), pts[1]],
        [None, pts[2], None, pts[2], pts[2]],
        [pts[3], pts[4], P.g1_add(pts[3], pts[4])],
        [None, None],
        [pts[0], pts[1], P.g1_neg(P.g1_add(pts[0], pts[1]))],
        [pts[5]] * 4,
        [P.G1_GEN_FIXED_TAU, P.G1_GEN_FIXED_TAU, P.G1_GEN],
    ]
    for c in cases:
        arr = np.concatenate([g1b(p) for p in c])
        exp = None
        for p in c:
            exp = P.g1_add(exp, p)
        o = np.zeros(24, dtype=np.uint32)
        L.h_g1_sum_madd(pp(arr), len(c), pp(o))
        assert g1t(o) == exp
        L.h_g1_sum_add(pp(arr), len(c), pp(o))
        assert g1t(o) == exp
    for k in [0, 1, 2, 3, P.R_MOD - 1, rng.randrange(P.R_MOD)]:
        o = np.zeros(24, dtype=np.uint32)
        L.h_g1_mul(pp(g1b(pts[0])), pp(u32(k, 8)), pp(o))
        assert g1t(o) == P.g1_mul(pts[0], k)
    o = np.zeros(24, dtype=np.uint32)
    L.h_g1_dbl_n(pp(g1b(pts[1])), 16, pp(o))
    assert g1t(o) == P.g1_mul(pts[1], 1 << 16)
