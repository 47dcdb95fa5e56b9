"""Known answers dumped from the REAL reference by scripts/pin/pin_against_reference.sh (needs cargo + ICICLE v3.8.0, so it
cannot run in the build image).  When tests/golden/reference_pins.json is present these tests turn "parity unpinned" into a
pinned oracle: roots of unity (the 5-based 2^32-th root is an inference until then), a bivariate NTT with and without
cosets, and an MSM.  Without the file they skip and say why."""
import json
import os

import pytest

import pyref as P

PINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_pins.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PINS), reason="tests/golden/reference_pins.json absent: run scripts/pin/pin_against_reference.sh "
                                "on a machine that builds the reference (parity stays unpinned until then)")


def _pins():
    return json.load(open(PINS))


def _ints(xs):
    return [int(v, 16) for v in xs]


def test_roots_of_unity_match_the_reference_domain():
    for k, v in _pins()["root_of_unity"].items():
        assert P.root_of_unity(1 << int(k)) == int(v, 16), f"omega_(2^{k})"


def test_bintt_matches_the_reference():
    g = _pins()["bintt"]
    x, y, a = g["x"], g["y"], _ints(g["in"])
    assert P.bintt(a, x, y) == _ints(g["fwd"])
    assert P.bintt(a, x, y, coset_x=int(g["coset_x"], 16), coset_y=int(g["coset_y"], 16)) == _ints(g["fwd_coset"])
    assert P.bintt(a, x, y, inverse=True) == _ints(g["inv"])


def test_msm_matches_the_reference():
    g = _pins()["msm"]
    ks, ss = _ints(g["base_multipliers"]), _ints(g["scalars"])
    bases = [P.g1_mul(P.G1_GEN, k) for k in ks]
    assert P.msm_g1(ss, bases) == (int(g["result"]["x"], 16), int(g["result"]["y"], 16))


@pytest.mark.gpu
def test_cuda_library_matches_the_reference_pins():
    import numpy as np

    import tokamak_b200 as T
    from util import frs, g1s, to_ints

    ctx = T.Context(0)
    ctx.init_ntt_domain_for_size(1 << 16)
    g = _pins()["bintt"]
    a = frs(_ints(g["in"]))
    assert to_ints(ctx.bintt_host(a, g["x"], g["y"], T.FORWARD)) == _ints(g["fwd"])
    assert to_ints(ctx.bintt_host(a, g["x"], g["y"], T.FORWARD, int(g["coset_x"], 16), int(g["coset_y"], 16))) == _ints(g["fwd_coset"])
    m = _pins()["msm"]
    bases = g1s([P.g1_mul(P.G1_GEN, k) for k in _ints(m["base_multipliers"])])
    got = ctx.msm_g1_host(frs(_ints(m["scalars"])), bases)
    exp = g1s([(int(m["result"]["x"], 16), int(m["result"]["y"], 16))])[0]
    assert np.array_equal(got, exp)
    ctx.close()
