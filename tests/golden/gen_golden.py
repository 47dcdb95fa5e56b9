"""Generates tests/golden/golden.json from the big-integer restatement oracle/pyref.py.

The reference ships no golden vectors for this path and cannot be built here, so
these vectors pin OUR restatement (two independent implementations -- pyref.py and oracle.c -- and the
CUDA path must all reproduce them).  Regenerate with:  python tests/golden/gen_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyref as P  # noqa: E402


def hx(v):
    return "0x%064x" % v


def pt(p):
    return None if p is None else ["0x%096x" % p[0], "0x%096x" % p[1]]


def main():
    g = {"constants": {
        "r": hx(P.R_MOD), "rou_2_32": hx(P.ROU),
        "roots_of_unity": {str(k): hx(P.root_of_unity(1 << k)) for k in (1, 2, 3, 8, 12, 20, 23, 32)},
        "g1_gen_fixed_tau": pt(P.G1_GEN_FIXED_TAU),
    }}
    rng = P.SplitMix64(20260418)
    # biNTT 8 x 4: plain, coset on both axes, inverse with coset
    x, y = 8, 4
    a = rng.frs(x * y)
    gx, gy = rng.fr(), rng.fr()
    g["bintt"] = {
        "x": x, "y": y, "in": [hx(v) for v in a], "coset_x": hx(gx), "coset_y": hx(gy),
        "fwd": [hx(v) for v in P.bintt(a, x, y)],
        "fwd_coset": [hx(v) for v in P.bintt(a, x, y, False, gx, gy)],
        "inv": [hx(v) for v in P.bintt(a, x, y, True)],
        "inv_coset": [hx(v) for v in P.bintt(a, x, y, True, gx, gy)],
    }
    # degenerate axes
    b = rng.frs(16)
    g["ntt_1d"] = {"in": [hx(v) for v in b], "fwd_x16": [hx(v) for v in P.bintt(b, 16, 1)], "inv_y16": [hx(v) for v in P.bintt(b, 1, 16, True)]}
    # polynomial ops on 8 x 4
    px, py = rng.fr(), rng.fr()
    qx, qy, r = P.div_by_ruffini(a, x, y, px, py)
    g["poly"] = {
        "point": [hx(px), hx(py)],
        "eval": hx(P.eval_xy(a, x, y, px, py)),
        "scale": [hx(v) for v in P.scale_coeffs(a, x, y, px, py)],
        "ruffini_qx": [hx(v) for v in qx], "ruffini_qy": [hx(v) for v in qy], "ruffini_r": hx(r),
    }
    c, d = 4, 2
    vqx, vqy = P.div_by_vanishing_opt(a, x, y, c, d)  # any polynomial: the recurrences are total functions
    g["vanishing"] = {"c": c, "d": d, "qx": [hx(v) for v in vqx], "qy": [hx(v) for v in vqy]}
    m, nx, ny = P.poly_mul(a, x, y, a, x, y)
    g["mul_self"] = {"nx": nx, "ny": ny, "out": [hx(v) for v in m]}
    # MSM: 12 points incl. identity, duplicate, negation, edge scalars
    ks = rng.frs(9)
    pts = [P.g1_mul(P.G1_GEN, k) for k in ks]
    pts += [None, pts[0], P.g1_neg(pts[1])]
    ss = rng.frs(8) + [0, 1, P.R_MOD - 1, 5]
    g["msm"] = {"scalars": [hx(s) for s in ss], "points": [pt(p) for p in pts], "result": pt(P.msm_naive(ss, pts))}
    # commitment of the 8x4 polynomial against a fixed-tau style CRS grid 8 x 4: xy_powers[4*h+i] = tau_x^h tau_y^i * G
    tx, ty = P.TAU_FIXED["x"], P.TAU_FIXED["y"]
    grid = [P.g1_mul(P.G1_GEN_FIXED_TAU, pow(tx, h, P.R_MOD) * pow(ty, i, P.R_MOD) % P.R_MOD) for h in range(8) for i in range(4)]
    com = P.encode_poly(a, x, y, grid, 8, 4)
    assert com == P.g1_mul(P.G1_GEN_FIXED_TAU, P.eval_xy(a, x, y, tx, ty))  # setup/trusted-setup/src/main.rs:222-246
    g["commit"] = {"grid": [pt(p) for p in grid], "result": pt(com)}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=0)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
