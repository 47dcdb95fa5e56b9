"""The C++ host mirror of the `libs` API (tokamak-zk-evm_b200/host/cpp): its host-side ScalarField arithmetic is checked
here without a GPU; the reference's library tests restated against it run on the GPU through the C-ABI."""
import os
import subprocess

import pytest

import pyref as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tokamak-zk-evm_b200", "host", "cpp")
LIB = os.path.join(ROOT, "tokamak-zk-evm_b200", "lib", "libtokamak_b200.so")


@pytest.fixture(scope="module")
def binary():
    if not os.path.exists(LIB):
        pytest.skip("libtokamak_b200.so is not built")
    subprocess.check_call(["make", "-s", "-C", CPP])
    return os.path.join(CPP, "test_libs")


def test_host_scalar_field_arithmetic(binary):
    out = subprocess.run([binary, "--scalar"], capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == 28
    R = P.R_MOD
    for i in range(0, len(out), 7):
        a, b, s, d, m, inv, pw = [int(x, 16) for x in out[i:i + 7]]
        assert s == (a + b) % R and d == (a - b) % R and m == a * b % R
        assert inv == pow(a, R - 2, R) and pw == pow(a, 65537, R)


@pytest.mark.gpu
def test_reference_library_tests_through_the_cpp_mirror(binary):
    """libs/src/tests.rs restated in C++ (test_from_evals, test_coset_ntt_matches_manual_scaling, test_mul_polynomial,
    test_div_by_ruffini, test_div_by_vanishing_opt_basic, MSM = scalar multiplication, encode_poly = P(tau) G, ...)."""
    r = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL PASSED" in r.stdout
