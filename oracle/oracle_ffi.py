"""ctypes loader for oracle/liboracle.so (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.c header).

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs only.
All buffers are numpy uint64 arrays in canonical little-endian form:
Fr -> shape (n, 4); G1 affine -> shape (n, 12), all-zero row = identity.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(ctypes.c_int(n))


def fr_from_int(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


def fr_to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes(), "little")


def frs_from_ints(vs):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vs), dtype=np.uint64).reshape(-1, 4).copy()


def frs_to_ints(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i : i + 32], "little") for i in range(0, len(b), 32)]


def g1_from_tuple(pt):
    if pt is None:
        return np.zeros(12, dtype=np.uint64)
    return np.frombuffer(pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little"), dtype=np.uint64).copy()


def g1_to_tuple(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    return None if (x == 0 and y == 0) else (x, y)


def fr_vec_op(op, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    lib().orc_fr_vec_op(ctypes.c_int({"add": 0, "sub": 1, "mul": 2}[op]), _p(a), _p(b), _p(out), ctypes.c_size_t(a.size // 4))
    return out


def fr_vec_inv(a):
    a = _u64(a)
    out = np.empty_like(a)
    lib().orc_fr_vec_inv(_p(a), _p(out), ctypes.c_size_t(a.size // 4))
    return out


def root_of_unity(log_n):
    out = np.empty(4, dtype=np.uint64)
    lib().orc_root_of_unity(ctypes.c_uint(log_n), _p(out))
    return out


def ntt(a, n, batch, columns_batch=False, inverse=False, coset=None):
    a = _u64(a)
    out = np.empty_like(a)
    c = _u64(coset) if coset is not None else None
    rc = lib().orc_ntt(_p(a), _p(out), ctypes.c_size_t(n), ctypes.c_size_t(batch), ctypes.c_int(columns_batch), ctypes.c_int(inverse), _p(c))
    assert rc == 0
    return out


def bintt(a, x, y, inverse=False, coset_x=None, coset_y=None):
    a = _u64(a)
    out = np.empty_like(a)
    cx = _u64(coset_x) if coset_x is not None else None
    cy = _u64(coset_y) if coset_y is not None else None
    rc = lib().orc_bintt(_p(a), _p(out), ctypes.c_size_t(x), ctypes.c_size_t(y), ctypes.c_int(inverse), _p(cx), _p(cy))
    assert rc == 0
    return out


def poly_mul_padded(a, b, x, y):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    rc = lib().orc_poly_mul_padded(_p(a), _p(b), _p(out), ctypes.c_size_t(x), ctypes.c_size_t(y))
    assert rc == 0
    return out


def scale_coeffs(a, x, y, sx=None, sy=None):
    a = _u64(a)
    out = np.empty_like(a)
    sx = _u64(sx) if sx is not None else None
    sy = _u64(sy) if sy is not None else None
    lib().orc_scale_coeffs(_p(a), _p(out), ctypes.c_size_t(x), ctypes.c_size_t(y), _p(sx), _p(sy))
    return out


def eval_xy(a, x, y, px, py):
    a, px, py = _u64(a), _u64(px), _u64(py)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_eval(_p(a), ctypes.c_size_t(x), ctypes.c_size_t(y), _p(px), _p(py), _p(out))
    return out


def div_by_vanishing_opt(p, x, y, c, d):
    p = _u64(p)
    qx = np.empty((x * y, 4), dtype=np.uint64)
    qy = np.empty((c * y, 4), dtype=np.uint64)
    rc = lib().orc_div_by_vanishing_opt(_p(p), ctypes.c_size_t(x), ctypes.c_size_t(y), ctypes.c_size_t(c), ctypes.c_size_t(d), _p(qx), _p(qy))
    assert rc == 0
    return qx, qy


def div_by_ruffini(p, x, y, px, py):
    p, px, py = _u64(p), _u64(px), _u64(py)
    qx = np.empty((x * y, 4), dtype=np.uint64)
    qy = np.empty((y, 4), dtype=np.uint64)
    r = np.empty(4, dtype=np.uint64)
    lib().orc_div_by_ruffini(_p(p), ctypes.c_size_t(x), ctypes.c_size_t(y), _p(px), _p(py), _p(qx), _p(qy), _p(r))
    return qx, qy, r


def g1_is_on_curve(pt):
    pt = _u64(pt)
    return bool(lib().orc_g1_is_on_curve(_p(pt)))


def g1_add(a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def g1_mul(pt, k):
    pt, k = _u64(pt), _u64(k)
    out = np.empty(12, dtype=np.uint64)
    lib().orc_g1_mul(_p(pt), _p(k), _p(out))
    return out


def g1_fixed_base_mul_batch(base, scalars):
    base, scalars = _u64(base), _u64(scalars)
    n = scalars.size // 4
    out = np.empty((n, 12), dtype=np.uint64)
    lib().orc_g1_fixed_base_mul_batch(_p(base), _p(scalars), ctypes.c_size_t(n), _p(out))
    return out


def msm_g1(scalars, bases, naive=False):
    scalars, bases = _u64(scalars), _u64(bases)
    n = scalars.size // 4
    assert bases.size // 12 == n
    out = np.empty(12, dtype=np.uint64)
    if naive:
        lib().orc_msm_g1_naive(_p(scalars), _p(bases), ctypes.c_size_t(n), _p(out))
    else:
        rc = lib().orc_msm_g1(_p(scalars), _p(bases), ctypes.c_size_t(n), _p(out))
        assert rc == 0
    return out


def msm_g1_rect(scalars, s_row_stride, bases, b_row_stride, rows, cols):
    scalars, bases = _u64(scalars), _u64(bases)
    out = np.empty(12, dtype=np.uint64)
    rc = lib().orc_msm_g1_rect(_p(scalars), ctypes.c_size_t(s_row_stride), _p(bases), ctypes.c_size_t(b_row_stride), ctypes.c_size_t(rows), ctypes.c_size_t(cols), _p(out))
    assert rc == 0
    return out


def fr_inner_product(a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty(4, dtype=np.uint64)
    lib().orc_fr_inner_product(_p(a), _p(b), ctypes.c_size_t(a.size // 4), _p(out))
    return out


def random_fr(seed, n):
    out = np.empty((n, 4), dtype=np.uint64)
    lib().orc_random_fr(ctypes.c_uint64(seed), ctypes.c_size_t(n), _p(out))
    return out
