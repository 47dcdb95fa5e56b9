"""CPU: the two oracle restatements (pyref.py big-int, oracle.c Montgomery/OpenMP) against each other,
against the committed golden vectors and against the in-tree constants of the reference."""
import numpy as np
import pytest

import oracle_ffi as O
import pyref as P
from util import frs, fr1, g1_tuple, g1s, golden, ints, pt_from_golden, to_ints


@pytest.fixture(scope="module", autouse=True)
def _build():
    O.build()


def test_constants():
    g = golden()["constants"]
    assert int(g["r"], 16) == P.R_MOD
    assert int(g["rou_2_32"], 16) == P.ROU == pow(5, (P.R_MOD - 1) >> 32, P.R_MOD)
    assert pow(P.ROU, 1 << 31, P.R_MOD) == P.R_MOD - 1  # primitive 2^32-th root
    for k, v in g["roots_of_unity"].items():
        assert int(v, 16) == P.root_of_unity(1 << int(k))
        assert O.fr_to_int(O.root_of_unity(int(k))) == int(v, 16)
    # setup/trusted-setup/src/main.rs:71-74: the --fixed-tau generator is a curve point
    assert P.g1_is_on_curve(P.G1_GEN_FIXED_TAU) and O.g1_is_on_curve(g1s([P.G1_GEN_FIXED_TAU])[0])
    assert P.g1_mul(P.G1_GEN, P.R_MOD - 1) == P.g1_neg(P.G1_GEN)  # prime-order subgroup


def test_golden_bintt_both_oracles():
    g = golden()["bintt"]
    x, y, a = g["x"], g["y"], ints(g["in"])
    gx, gy = int(g["coset_x"], 16), int(g["coset_y"], 16)
    cases = {"fwd": (False, None, None), "fwd_coset": (False, gx, gy), "inv": (True, None, None), "inv_coset": (True, gx, gy)}
    for name, (inv, cx, cy) in cases.items():
        exp = ints(g[name])
        assert P.bintt(a, x, y, inv, cx, cy) == exp
        got = O.bintt(frs(a), x, y, inv, None if cx is None else fr1(cx), None if cy is None else fr1(cy))
        assert to_ints(got) == exp
    g1 = golden()["ntt_1d"]
    b = ints(g1["in"])
    assert to_ints(O.bintt(frs(b), 16, 1)) == ints(g1["fwd_x16"])
    assert to_ints(O.bintt(frs(b), 1, 16, True)) == ints(g1["inv_y16"])


def test_golden_poly_ops():
    g = golden()
    x, y, a = g["bintt"]["x"], g["bintt"]["y"], ints(g["bintt"]["in"])
    px, py = ints(g["poly"]["point"])
    assert O.fr_to_int(O.eval_xy(frs(a), x, y, fr1(px), fr1(py))) == int(g["poly"]["eval"], 16)
    assert to_ints(O.scale_coeffs(frs(a), x, y, fr1(px), fr1(py))) == ints(g["poly"]["scale"])
    qx, qy, r = O.div_by_ruffini(frs(a), x, y, fr1(px), fr1(py))
    assert to_ints(qx) == ints(g["poly"]["ruffini_qx"]) and to_ints(qy) == ints(g["poly"]["ruffini_qy"])
    assert O.fr_to_int(r) == int(g["poly"]["ruffini_r"], 16)
    v = g["vanishing"]
    vqx, vqy = O.div_by_vanishing_opt(frs(a), x, y, v["c"], v["d"])
    assert to_ints(vqx) == ints(v["qx"]) and to_ints(vqy) == ints(v["qy"])
    m = g["mul_self"]
    pa, nx, ny = P.resize(a, x, y, m["nx"], m["ny"])
    assert to_ints(O.poly_mul_padded(frs(pa), frs(pa), nx, ny)) == ints(m["out"])


def test_golden_msm_and_commit():
    g = golden()
    ss = ints(g["msm"]["scalars"])
    pts = [pt_from_golden(p) for p in g["msm"]["points"]]
    exp = pt_from_golden(g["msm"]["result"])
    assert P.msm_naive(ss, pts) == exp
    assert g1_tuple(O.msm_g1(frs(ss), g1s(pts))) == exp
    assert g1_tuple(O.msm_g1(frs(ss), g1s(pts), naive=True)) == exp
    grid = [pt_from_golden(p) for p in g["commit"]["grid"]]
    a = ints(g["bintt"]["in"])
    exp = pt_from_golden(g["commit"]["result"])
    assert g1_tuple(O.msm_g1_rect(frs(a), 4, g1s(grid), 4, 8, 4)) == exp
    # encode_poly(P) == P(tau_x, tau_y) * G  (setup/trusted-setup/src/main.rs:222-246)
    assert exp == P.g1_mul(P.G1_GEN_FIXED_TAU, P.eval_xy(a, 8, 4, P.TAU_FIXED["x"], P.TAU_FIXED["y"]))


def test_c_oracle_vs_pyref_random():
    a, b = O.random_fr(1, 64), O.random_fr(2, 64)
    assert to_ints(O.random_fr(7, 50)) == P.random_fr(7, 50)
    ai, bi = to_ints(a), to_ints(b)
    assert to_ints(O.fr_vec_op("mul", a, b)) == [u * v % P.R_MOD for u, v in zip(ai, bi)]
    assert to_ints(O.fr_vec_op("add", a, b)) == [(u + v) % P.R_MOD for u, v in zip(ai, bi)]
    assert to_ints(O.fr_vec_op("sub", a, b)) == [(u - v) % P.R_MOD for u, v in zip(ai, bi)]
    assert to_ints(O.fr_vec_inv(a)) == [P.fr_inv(u) for u in ai]
    x, y = 16, 8
    m = O.random_fr(3, x * y)
    mi = to_ints(m)
    cx, cy = O.random_fr(4, 1)[0], O.random_fr(5, 1)[0]
    for inv in (False, True):
        for gx, gy in ((None, None), (cx, None), (None, cy), (cx, cy)):
            exp = P.bintt(mi, x, y, inv, None if gx is None else O.fr_to_int(gx), None if gy is None else O.fr_to_int(gy))
            assert to_ints(O.bintt(m, x, y, inv, gx, gy)) == exp
    # column batch == row batch on the transpose (libs/src/tests.rs:519-588)
    cols = to_ints(O.ntt(m, x, y, columns_batch=True))
    for j in range(y):
        assert cols[j::y] == P.ntt(mi[j::y])


def test_oracle_msm_pippenger_vs_naive():
    G = g1s([P.G1_GEN])[0]
    ks = O.random_fr(9, 300)
    pts = O.g1_fixed_base_mul_batch(G, ks)
    for i in range(3):
        assert g1_tuple(pts[i]) == P.g1_mul(P.G1_GEN, O.fr_to_int(ks[i]))
    ss = O.random_fr(10, 300)
    exp = P.g1_mul(P.G1_GEN, O.fr_to_int(O.fr_inner_product(ss, ks)))
    assert g1_tuple(O.msm_g1(ss, pts)) == exp
    assert g1_tuple(O.msm_g1(ss[:40], pts[:40], naive=True)) == g1_tuple(O.msm_g1(ss[:40], pts[:40]))
    # all-equal bases (the setup-side generator MSM, iotools/mod.rs:1113-1135) and all-equal scalars
    same = np.tile(pts[0], (64, 1))
    assert g1_tuple(O.msm_g1(ss[:64], same)) == P.g1_mul(g1_tuple(pts[0]), sum(to_ints(ss[:64])) % P.R_MOD)


def test_reference_identities_on_oracle():
    """The identities libs/src/tests.rs pins: round trip (:107-131), coset == manual scaling (:134-180),
    product on the omega grid (:1042-1088), vanishing / Ruffini reconstruction (:935-952,1090-1237)."""
    rng = P.SplitMix64(5)
    x, y = 8, 8
    a = rng.frs(x * y)
    assert P.bintt(P.bintt(a, x, y), x, y, True) == a
    gx, gy = rng.fr(), rng.fr()
    assert P.bintt(a, x, y, False, gx, gy) == P.bintt(P.scale_coeffs(a, x, y, gx, gy), x, y)
    ev = P.bintt(a, x, y)
    wx, wy = P.root_of_unity(x), P.root_of_unity(y)
    assert ev[3 * y + 5] == P.eval_xy(a, x, y, pow(wx, 3, P.R_MOD), pow(wy, 5, P.R_MOD))
    # P = Qx (X^c - 1) + Qy (Y^d - 1)
    c, d = 4, 2
    qx0 = [rng.fr() if i < x - c else 0 for i in range(x) for j in range(y)]
    qy0 = [rng.fr() if j < y - d else 0 for i in range(c) for j in range(y)]
    Pm = [0] * (x * y)
    for i in range(x):
        for j in range(y):
            v = qx0[i * y + j]
            if v:
                Pm[(i + c) * y + j] = (Pm[(i + c) * y + j] + v) % P.R_MOD
                Pm[i * y + j] = (Pm[i * y + j] - v) % P.R_MOD
    for i in range(c):
        for j in range(y):
            v = qy0[i * y + j]
            if v:
                Pm[i * y + j + d] = (Pm[i * y + j + d] + v) % P.R_MOD
                Pm[i * y + j] = (Pm[i * y + j] - v) % P.R_MOD
    qx, qy = P.div_by_vanishing_opt(Pm, x, y, c, d)
    u, v = rng.fr(), rng.fr()
    assert P.eval_xy(Pm, x, y, u, v) == (P.eval_xy(qx, x, y, u, v) * (pow(u, c, P.R_MOD) - 1) + P.eval_xy(qy, c, y, u, v) * (pow(v, d, P.R_MOD) - 1)) % P.R_MOD
    rqx, rqy, r = P.div_by_ruffini(a, x, y, u, v)
    s, t = rng.fr(), rng.fr()
    assert P.eval_xy(a, x, y, s, t) == (P.eval_xy(rqx, x, y, s, t) * (s - u) + P.eval_xy(rqy, 1, y, s, t) * (t - v) + r) % P.R_MOD
