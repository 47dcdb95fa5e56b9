"""Host-slice vector helpers with the reference's names (libs/src/vector_operations/mod.rs:19-693) over the C-ABI.
Inputs/outputs are numpy uint64 arrays of shape (n, 4) in canonical little-endian form (or lists of ints)."""
import ctypes

import numpy as np

from . import FORWARD, INVERSE, OP_ADD, OP_DIV, OP_MUL, OP_SUB, R_MOD, DensePolynomialExt, _as_fr_array, _vp, check, fr_bytes, fr_to_int, frs_from_ints


def _binary(ctx, op, lhs, rhs):
    lhs, rhs = _as_fr_array(lhs), _as_fr_array(rhs)
    if lhs.shape != rhs.shape:
        raise ValueError("Mismatch of sizes of vectors to be pointwise-multiplied")
    return ctx.vec_op_host(op, lhs, rhs)


def point_mul_two_vecs(ctx, lhs, rhs):
    """:30-44"""
    return _binary(ctx, OP_MUL, lhs, rhs)


def point_div_two_vecs(ctx, numer, denom):
    """:46-56 (used by prove1, prove/src/lib.rs:1860)"""
    return _binary(ctx, OP_DIV, numer, denom)


def point_add_two_vecs(ctx, lhs, rhs):
    """:58-66"""
    return _binary(ctx, OP_ADD, lhs, rhs)


def _with_device(ctx, a, fn):
    a = _as_fr_array(a)
    d = ctx.upload_fr(a)
    try:
        return fn(d, a.shape[0])
    finally:
        ctx.dev_free(d)


def scale_vec(ctx, scaler, vec):
    """:68-80: scaler * vec"""
    def run(d, n):
        k, p = fr_bytes(scaler)
        check(ctx.lib.tkm_fr_vec_scale(ctx.h, p, ctypes.c_void_p(d), ctypes.c_void_p(d), n))
        return ctx.download_fr(d, n)
    return _with_device(ctx, vec, run)


def scalar_vec_add(ctx, scalar, vec):
    """:94-104: scalar + vec[i]"""
    v = _as_fr_array(vec)
    return _binary(ctx, OP_ADD, frs_from_ints([int(scalar)] * v.shape[0]), v)


def scalar_vec_sub(ctx, scalar, vec):
    """:82-92: scalar - vec[i]"""
    v = _as_fr_array(vec)
    return _binary(ctx, OP_SUB, frs_from_ints([int(scalar)] * v.shape[0]), v)


def inner_product_two_vecs(ctx, lhs, rhs):
    """:106-141: sum_k lhs[k] * rhs[k]"""
    a, b = _as_fr_array(lhs), _as_fr_array(rhs)
    if a.shape != b.shape:
        raise ValueError("Mismatch of sizes of vectors to be inner-producted")
    da, db = ctx.upload_fr(a), ctx.upload_fr(b)
    out = np.zeros(4, dtype=np.uint64)
    try:
        check(ctx.lib.tkm_fr_vec_reduce(ctx.h, 2, ctypes.c_void_p(da), ctypes.c_void_p(db), a.shape[0], _vp(out)))
    finally:
        ctx.dev_free(da)
        ctx.dev_free(db)
    return fr_to_int(out)


def vec_sum(ctx, vec):
    return _with_device(ctx, vec, lambda d, n: _reduce(ctx, 0, d, n))


def vec_product(ctx, vec):
    return _with_device(ctx, vec, lambda d, n: _reduce(ctx, 1, d, n))


def _reduce(ctx, op, d, n):
    out = np.zeros(4, dtype=np.uint64)
    check(ctx.lib.tkm_fr_vec_reduce(ctx.h, op, ctypes.c_void_p(d), None, n, _vp(out)))
    return fr_to_int(out)


def outer_product_two_vecs(ctx, col_vec, row_vec):
    """:551-581: res[i*cols + j] = col_vec[i] * row_vec[j]"""
    c, r = _as_fr_array(col_vec), _as_fr_array(row_vec)
    dc, dr = ctx.upload_fr(c), ctx.upload_fr(r)
    do = ctx.dev_alloc(c.shape[0] * r.shape[0] * 32)
    try:
        check(ctx.lib.tkm_fr_outer_product(ctx.h, ctypes.c_void_p(dc), ctypes.c_void_p(dr), ctypes.c_void_p(do), c.shape[0], r.shape[0]))
        return ctx.download_fr(do, c.shape[0] * r.shape[0])
    finally:
        for p in (dc, dr, do):
            ctx.dev_free(p)


def transpose_inplace(ctx, vec, row_size, col_size):
    """:143-170: a row_size x col_size row-major matrix becomes col_size x row_size."""
    a = _as_fr_array(vec)
    if a.shape[0] != row_size * col_size:
        raise ValueError("Error in transpose")
    d = ctx.upload_fr(a, to_mont=False)
    o = ctx.dev_alloc(a.nbytes)
    try:
        ctx.transpose_dev(d, o, row_size, col_size)
        return ctx.download_fr(o, a.shape[0], from_mont=False)
    finally:
        ctx.dev_free(d)
        ctx.dev_free(o)


def gen_evaled_lagrange_bases(ctx, val, size):
    """:19-28: coefficients of the polynomial whose values on the size-th roots of unity are val^i, i.e. the Lagrange
    basis polynomials evaluated at val (up to the 1/size the inverse transform carries)."""
    pows, acc = [], 1
    for _ in range(size):
        pows.append(acc)
        acc = acc * int(val) % R_MOD
    p = DensePolynomialExt.from_rou_evals(ctx, frs_from_ints(pows), size, 1)
    return p.copy_coeffs()


def resize(mat, curr_row_size, curr_col_size, target_row_size, target_col_size, zero=0):
    """:653-672: exact crop / zero-pad of a row-major matrix (host only, no arithmetic)."""
    a = _as_fr_array(mat).reshape(curr_row_size, curr_col_size, 4)
    out = np.zeros((target_row_size, target_col_size, 4), dtype=np.uint64)
    if zero:
        out[:] = frs_from_ints([zero])[0]
    r, c = min(curr_row_size, target_row_size), min(curr_col_size, target_col_size)
    out[:r, :c] = a[:r, :c]
    return out.reshape(-1, 4)


def extend_monomial_vec(ctx, mono_vec, target_len):
    """:639-651: extend [1, x, x^2, ...] to target_len entries (or truncate)."""
    v = _as_fr_array(mono_vec)
    n = v.shape[0]
    if target_len <= n:
        return v[:target_len].copy()
    ints = [fr_to_int(v[i]) for i in range(n)]
    x = ints[1]
    while len(ints) < target_len:
        ints.append(ints[-1] * x % R_MOD)
    return frs_from_ints(ints)
