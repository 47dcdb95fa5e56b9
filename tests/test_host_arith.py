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


LAMBDA = 0xAC45A4010001A40200000000FFFFFFFF


def test_glv_split_and_digits(L):
    """csrc/glv.cuh: k = k1 + k2*lambda (mod r) with both halves below 2^127, and the signed-digit streams k_decompose
    emits (plain and GLV) rebuild the scalar for every window width the MSM picks."""
    assert (LAMBDA * LAMBDA + LAMBDA + 1) == P.R_MOD
    rng = random.Random(11)
    r = P.R_MOD
    edge = [0, 1, 2, r - 1, r - 2, LAMBDA - 1, LAMBDA, LAMBDA + 1, LAMBDA // 2, LAMBDA // 2 + 1, (LAMBDA + 1) // 2 * LAMBDA,
            ((LAMBDA + 1) // 2 + 1) * LAMBDA, ((LAMBDA + 1) // 2 + 1) * LAMBDA - 1, (LAMBDA + 1) * LAMBDA, r // 2, r // 2 + 1,
            (1 << 127) - 1, 1 << 127, (1 << 128) - 1, 1 << 128, (1 << 255) % r, (1 << 64) - 1,
            r, r + 1, 2 * r, 2 * r + 5, (1 << 256) - 1]  # non-canonical inputs fold into [0, r)
    vals = edge + [rng.randrange(r) for _ in range(3000)] + [rng.randrange(1 << k) for k in (8, 32, 64, 126, 127, 128, 129, 200) for _ in range(40)]
    vals += [q * LAMBDA + d for q in (0, 1, (LAMBDA + 1) // 2 - 1, (LAMBDA + 1) // 2, (LAMBDA + 1) // 2 + 1, LAMBDA) for d in (0, 1, LAMBDA // 2, LAMBDA // 2 + 1, LAMBDA - 1)]
    out = np.zeros(10, dtype=np.uint32)
    digs = np.zeros(160, dtype=np.int32)
    for k in vals:
        L.h_glv_split(pp(u32(k, 8)), pp(out))
        m1, m2 = toint(out[0:4]), toint(out[4:8])
        k1 = -m1 if out[8] else m1
        k2 = -m2 if out[9] else m2
        assert m1 < (1 << 127) and m2 < (1 << 127), hex(k)
        assert (k1 + k2 * LAMBDA - k) % r == 0, hex(k)
        for c in (4, 7, 8, 11, 13, 15, 16, 17, 20):
            nw = L.h_msm_digits(pp(u32(k, 8)), c, 1, pp(digs))
            wh = (128 + c - 1) // c
            assert nw == 2 * wh
            d = [int(x) for x in digs[:nw]]
            assert all(abs(x) <= (1 << (c - 1)) for x in d)  # bucket index |digit| - 1 < 2^(c-1)
            assert sum(x << (c * w) for w, x in enumerate(d[:wh])) == k1
            assert sum(x << (c * w) for w, x in enumerate(d[wh:])) == k2
            if k < (1 << 255):
                nw = L.h_msm_digits(pp(u32(k, 8)), c, 0, pp(digs))
                assert nw == (256 + c - 1) // c
                assert sum(int(x) << (c * w) for w, x in enumerate(digs[:nw])) == k


def test_glv_endomorphism(L):
    rng = random.Random(12)
    for _ in range(4):
        pt = P.g1_mul(P.G1_GEN, rng.randrange(P.R_MOD))
        o = np.zeros(24, dtype=np.uint32)
        L.h_g1_phi(pp(g1b(pt)), pp(o))
        assert g1t(o) == P.g1_mul(pt, LAMBDA)


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_binary_gcd_inverse(L, name):
    """Fp::inv_bgcd (binary extended Euclid, used by the MSM tail) equals the Fermat inverse; inv(0) = 0."""
    mod, n, fn = (P.R_MOD, 8, L.h_fr_op) if name == "fr" else (P.Q_MOD, 12, L.h_fq_op)
    rng = random.Random(21)
    R = 1 << (32 * n)
    vals = [0, 1, 2, 3, 4, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1, R % mod, pow(R, -1, mod), (1 << 32) - 1, 1 << 32, 1 << (32 * n - 2)]
    vals += [rng.randrange(mod) for _ in range(300)] + [1 << k for k in range(0, 32 * n - 1, 37)]
    for a in vals:
        a %= mod
        o = np.zeros(n, dtype=np.uint32)
        fn(7, pp(u32(a, n)), pp(u32(a, n)), pp(o))
        assert toint(o) == (pow(a, mod - 2, mod) if a else 0), (name, hex(a))


@pytest.mark.parametrize("name", ["fr", "fq"])
def test_binary_gcd_on_approximations_inverse(L, name):
    """Fp::inv_fast (Pornin's binary GCD on 64-bit approximations, checked, with the inv_bgcd fallback) equals the Fermat
    inverse on edge values, powers of two, values around limb boundaries and random elements; for Fq also inv_pornin ALONE
    (op 9 returns 0 when its own (a, b) = (0, 1) check fails), so the fast path itself is what is tested, not the fallback."""
    mod, n, fn = (P.R_MOD, 8, L.h_fr_op) if name == "fr" else (P.Q_MOD, 12, L.h_fq_op)
    rng = random.Random(22)
    R = 1 << (32 * n)
    vals = [0, 1, 2, 3, 4, mod - 1, mod - 2, mod >> 1, (mod >> 1) + 1, R % mod, pow(R, -1, mod), (1 << 32) - 1, 1 << 32, 1 << (32 * n - 2)]
    vals += [rng.randrange(mod) for _ in range(3000)] + [1 << k for k in range(0, 32 * n - 1, 7)]
    vals += [((1 << k) - 1) % mod for k in range(1, 32 * n, 11)] + [(mod - (1 << k)) % mod for k in range(0, 32 * n - 2, 13)]
    vals += [rng.randrange(1 << k) for k in (8, 31, 32, 33, 63, 64, 65, 96, 127, 200) for _ in range(30)]
    rinv = pow(R, -1, mod)
    vals += [x * rinv % mod for x in (1, 2, 3, (1 << 31), (1 << 32) - 1, (1 << 33) + 1, (1 << 64) - 1, 1 << 95)]  # tiny Montgomery forms: long runs of zero limbs
    for a in vals:
        a %= mod
        exp = pow(a, mod - 2, mod) if a else 0
        o = np.zeros(n, dtype=np.uint32)
        fn(8, pp(u32(a, n)), pp(u32(a, n)), pp(o))
        assert toint(o) == exp, (name, hex(a))
        if name == "fq" and a:
            fn(9, pp(u32(a, n)), pp(u32(a, n)), pp(o))
            assert toint(o) == exp, ("inv_pornin alone", hex(a))


def test_fq_dot2(L):
    """Fq::dot2 = a*b + c*d under one interleaved Montgomery reduction (the Y3 of the point formulas), raw limbs."""
    q, n = P.Q_MOD, 12
    rng = random.Random(31)
    rinv = pow(1 << 384, -1, q)
    edge = [0, 1, q - 1, q - 2, (1 << 384) % q, (1 << 32) - 1, q >> 1, int("ffffffff" * 11, 16)]
    quads = [(a, b, c, d) for a in edge for b in edge[:4] for c in edge[1:5] for d in edge[2:6]]
    quads += [(q - 1, q - 1, q - 1, q - 1), (0, 0, 0, 0), (q - 1, q - 1, 0, 0), (0, 5, q - 1, q - 1)]
    quads += [tuple(rng.randrange(q) for _ in range(4)) for _ in range(2000)]
    for a, b, c, d in quads:
        o = np.zeros(n, dtype=np.uint32)
        L.h_fq_dot2(pp(u32(a, n)), pp(u32(b, n)), pp(u32(c, n)), pp(u32(d, n)), pp(o))
        assert toint(o) == (a * b + c * d) * rinv % q, (hex(a), hex(b), hex(c), hex(d))
