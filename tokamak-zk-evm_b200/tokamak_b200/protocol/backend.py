"""The product backend of the protocol driver: every polynomial, commitment and sparse MSM runs on the B200 through the
C-ABI of libtokamak_b200 (no CPU fallback: constructing it without a GPU fails in tkm_ctx_create).

The driver talks to a backend through this small surface (the CPU twin used by the tests lives with the test oracle,
outside this package, and implements the same methods):
  from_coeffs / from_rou_evals -> polynomial with  + - * (poly | int), mul_monomial, scale_coeffs_x/y, eval,
                                   div_by_vanishing_opt, div_by_ruffini, to_rou_evals, clone
  make_table(col, row, base)    -> G1 table  T[j][i] = col[j] * row[i] * base   (Sigma1 components)
  commit(table, poly)           -> sum_ij c_ij T[i][j]                          (Sigma1::encode_poly)
  msm_indexed(table, idx, s)    -> sum_k s_k T.flat[idx_k]                      (msm_g1_bases over gathered rows)
  g1_add / g1_sub / g1_mul, recursion_evals
G1 points cross this interface as (x, y) integer tuples, None = identity."""
import ctypes

import numpy as np

from .. import DensePolynomialExt, _as_fr_array, _vp, check, frs_from_ints
from .fr import R_MOD


def g1_to_tuple(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).reshape(12).tobytes()
    x, y = int.from_bytes(b[:48], "little"), int.from_bytes(b[48:], "little")
    return None if x == 0 and y == 0 else (x, y)


def g1_from_tuple(p):
    if p is None:
        return np.zeros(12, dtype=np.uint64)
    return np.frombuffer(p[0].to_bytes(48, "little") + p[1].to_bytes(48, "little"), dtype=np.uint64).copy()


class GpuTable:
    def __init__(self, ctx, handle, rows, cols):
        self.ctx, self.h, self.rows, self.cols = ctx, handle, rows, cols

    def device_ptr(self):
        p = ctypes.c_void_p()
        check(self.ctx.lib.tkm_crs_device_ptr(self.h, ctypes.byref(p), None, None))
        return p.value

    def precompute(self, window_bits=20):
        """Fixed-base tables 2^(c w) P for every point (serving mode: one CRS reused across proofs)."""
        check(self.ctx.lib.tkm_crs_precompute(self.ctx.h, self.h, window_bits))

    def points_host(self):
        """Canonical affine points (rows*cols, 12) -- test/debug only."""
        n = self.rows * self.cols
        tmp = self.ctx.dev_alloc(n * 96)
        check(self.ctx.lib.tkm_g1_bases_from_mont(self.ctx.h, self.device_ptr(), tmp, n))
        out = np.empty((n, 12), dtype=np.uint64)
        self.ctx.d2h(out, tmp)
        self.ctx.dev_free(tmp)
        return out

    def mont_bytes_host(self):
        """The table exactly as it sits on the device: x || y, 48-byte little-endian Montgomery coordinates -- the
        `ffjs-g1-affine-96` encoding of the TZBWASM1 CRS container (protocol/crs_io.py)."""
        n = self.rows * self.cols
        out = np.empty((n, 12), dtype=np.uint64)
        self.ctx.d2h(out, self.device_ptr())
        return out

    def close(self):
        if self.h:
            self.ctx.lib.tkm_crs_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _PendingCommit:
    def __init__(self, ctx, ticket):
        self.ctx, self.ticket, self.value, self.done = ctx, ticket, None, False

    def get(self):
        if not self.done:
            out = np.zeros(12, dtype=np.uint64)
            check(self.ctx.lib.tkm_commit_end(self.ctx.h, self.ticket, _vp(out)))
            self.value, self.done = g1_to_tuple(out), True
        return self.value


class GpuBackend:
    name = "b200"

    def __init__(self, ctx):
        self.ctx = ctx

    # ---- polynomials
    def from_coeffs(self, coeffs, x_size, y_size):
        return DensePolynomialExt.from_coeffs(self.ctx, coeffs, x_size, y_size)

    def from_rou_evals(self, evals, x_size, y_size):
        return DensePolynomialExt.from_rou_evals(self.ctx, evals, x_size, y_size)

    def init_ntt_domain(self, size):
        self.ctx.init_ntt_domain_for_size(size)

    def reserve(self, nbytes):
        """Grow the stream-ordered memory pool once (the pool never returns memory to the driver): later polynomial and
        MSM scratch allocations are then served without driver calls.  A prove at the reference shape peaks at a few GB; when
        the pool has to grow inside a stage that stage stalls for ~0.7 s."""
        p = self.ctx.dev_alloc(nbytes)
        self.ctx.dev_free(p)
        self.ctx.sync()

    def uvw_polys(self, params, csr, wt):
        """read_R1CS_gen_uvwXY on the device (tkm_r1cs_uvw_polys): sparse R1CS x witness + three inverse biNTTs."""
        u, v, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        vals = np.ascontiguousarray(wt.values)
        check(self.ctx.lib.tkm_r1cs_uvw_polys(self.ctx.h, len(csr.n_rows), _vp(csr.n_rows), _vp(csr.rp_base), _vp(csr.row_ptr), csr.row_ptr.shape[0],
                                              _vp(csr.wire), _vp(csr.coeff), csr.wire.shape[0], _vp(wt.sub_of_col), _vp(wt.var_off), _vp(vals),
                                              vals.shape[0], params.n, params.s_max, ctypes.byref(u), ctypes.byref(v), ctypes.byref(w)))
        return DensePolynomialExt(self.ctx, u), DensePolynomialExt(self.ctx, v), DensePolynomialExt(self.ctx, w)

    # ---- G1 tables (Sigma1 components), built and kept on the device
    def make_table(self, col, row, base):
        ctx = self.ctx
        rows, cols = len(col), len(row)
        n = rows * cols
        d_col = ctx.upload_fr(frs_from_ints(col))
        d_row = ctx.upload_fr(frs_from_ints(row))
        d_s = ctx.dev_alloc(n * 32)
        check(ctx.lib.tkm_fr_outer_product(ctx.h, d_col, d_row, d_s, rows, cols))
        d_pts = ctx.dev_alloc(n * 96)
        b = g1_from_tuple(base)
        check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, _vp(b), d_s, 1, n, d_pts))
        for p in (d_col, d_row, d_s):
            ctx.dev_free(p)
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_crs_from_device(ctx.h, d_pts, rows, cols, 1, ctypes.byref(h)))
        return GpuTable(ctx, h, rows, cols)

    def table_from_points(self, points, rows, cols):
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_crs_upload(self.ctx.h, _vp(pts), rows, cols, ctypes.byref(h)))
        return GpuTable(self.ctx, h, rows, cols)

    def table_from_mont_bytes(self, data, rows, cols):
        """Upload points that are already in the device layout (Montgomery little-endian, e.g. a TZBWASM1 CRS section read
        through a memory map): one H2D copy, no conversion."""
        arr = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data)
        if arr.nbytes != rows * cols * 96:
            raise ValueError("table size mismatch")
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_crs_upload_mont(self.ctx.h, _vp(arr), rows, cols, ctypes.byref(h)))
        return GpuTable(self.ctx, h, rows, cols)

    def commit(self, table, poly):
        out = np.zeros(12, dtype=np.uint64)
        check(self.ctx.lib.tkm_poly_commit(self.ctx.h, poly.h, table.h, _vp(out)))
        return g1_to_tuple(out)

    def commit_async(self, table, poly):
        """tkm_poly_commit_begin: queue the commitment and return at once; .get() (tkm_commit_end) waits for its tail.
        The recombination tail of this MSM overlaps whatever is queued next on the context stream."""
        t = ctypes.c_int32()
        check(self.ctx.lib.tkm_poly_commit_begin(self.ctx.h, poly.h, table.h, ctypes.byref(t)))
        return _PendingCommit(self.ctx, t.value)

    def msm_indexed(self, table, idx, scalars):
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        s = _as_fr_array(scalars)
        n = idx.shape[0]
        if n == 0:
            return None
        assert s.shape[0] == n
        ctx = self.ctx
        d_s = ctx.upload_fr(s, to_mont=False)
        d_i = ctx.dev_alloc(n * 4)
        ctx.h2d(d_i, idx)
        out = ctx.msm_g1_indexed_dev(d_s, False, table.device_ptr(), d_i, n)
        ctx.dev_free(d_s)
        ctx.dev_free(d_i)
        return g1_to_tuple(out)

    def msm_indexed_async(self, table, idx, scalars):
        """tkm_msm_g1_indexed_begin: the same sparse-gather MSM queued; .get() resolves it (tkm_commit_end)."""
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        s = _as_fr_array(scalars)
        n = idx.shape[0]
        if n == 0:
            return None
        assert s.shape[0] == n
        ctx = self.ctx
        d_s = ctx.upload_fr(s, to_mont=False)
        d_i = ctx.dev_alloc(n * 4)
        ctx.h2d(d_i, idx)
        t = ctypes.c_int32()
        check(ctx.lib.tkm_msm_g1_indexed_begin(ctx.h, d_s, 0, table.device_ptr(), d_i, n, ctypes.byref(t)))
        ctx.dev_free(d_s)  # stream-ordered: released after the queued kernels have read them
        ctx.dev_free(d_i)
        return _PendingCommit(ctx, t.value)

    # ---- placement variables kept on the device for the duration of Prover.init
    def witness_device(self, wt):
        """The witness table (canonical limbs) uploaded once per WitnessTable; released by release_witness."""
        dev = getattr(wt, "_dev", None)
        if dev is None:
            dev = self.ctx.upload_fr(wt.values, to_mont=False)
            wt._dev = dev
        return dev

    def release_witness(self, wt):
        dev = getattr(wt, "_dev", None)
        if dev is not None:
            self.ctx.dev_free(dev)
            wt._dev = None

    def msm_indexed_witness(self, table, idx, wt, rows, defer=False):
        """sum_k values[rows[k]] * table[idx[k]]: the scalars are gathered on the device (tkm_fr_gather) from the resident
        witness table, so only the two u32 index vectors cross PCIe."""
        n = idx.shape[0]
        if n == 0:
            return None
        ctx = self.ctx
        d_vals = self.witness_device(wt)
        d_i, d_r, d_s = ctx.dev_alloc(n * 4), ctx.dev_alloc(n * 4), ctx.dev_alloc(n * 32)
        ctx.h2d(d_i, idx)
        ctx.h2d(d_r, rows)
        check(ctx.lib.tkm_fr_gather(ctx.h, d_vals, wt.values.shape[0], d_r, n, d_s))
        t = ctypes.c_int32()
        check(ctx.lib.tkm_msm_g1_indexed_begin(ctx.h, d_s, 0, table.device_ptr(), d_i, n, ctypes.byref(t)))
        for p_ in (d_i, d_r, d_s):
            ctx.dev_free(p_)  # stream-ordered
        pend = _PendingCommit(ctx, t.value)
        return pend if defer else pend.get()

    def interface_poly(self, params, wt):
        """gen_bXY (polynomial_structures/mod.rs:132-162) on the device: the interface wires' values scattered from the
        resident witness table into an m_I x s_max evaluation table, then one inverse biNTT.  Nothing of that size is
        built on the host."""
        ctx, lib = self.ctx, self.ctx.lib
        m_i, s_max = params.l_D - params.l, params.s_max
        n = m_i * s_max
        idx, rows = wt.gather_indices(params.l, params.l_D, s_max)
        k = idx.shape[0]
        d_vals = self.witness_device(wt)
        d_ev = ctx.dev_alloc(n * 32)
        from .. import fr_bytes

        check(lib.tkm_fr_vec_fill(ctx.h, fr_bytes(0)[1], d_ev, n))
        if k:
            d_i, d_r = ctx.dev_alloc(k * 4), ctx.dev_alloc(k * 4)
            ctx.h2d(d_i, idx)
            ctx.h2d(d_r, rows)
            check(lib.tkm_fr_scatter_from_table(ctx.h, d_ev, n, d_i, d_vals, wt.values.shape[0], d_r, k))
            ctx.dev_free(d_i)
            ctx.dev_free(d_r)
        check(lib.tkm_fr_to_mont(ctx.h, d_ev, d_ev, n))
        h = ctypes.c_void_p()
        check(lib.tkm_poly_from_device(ctx.h, d_ev, m_i, s_max, ctypes.byref(h)))
        ctx.dev_free(d_ev)
        q = DensePolynomialExt(ctx, h)
        check(lib.tkm_poly_ntt_inplace(ctx.h, q.h, 1, None, None))
        return q

    def msm_points_async(self, points, scalars):
        """msm_points queued (tkm_msm_g1_begin over freshly uploaded points)."""
        ctx = self.ctx
        pts = np.ascontiguousarray(np.stack([g1_from_tuple(p) for p in points]), dtype=np.uint64)
        n = pts.shape[0]
        d_p = ctx.dev_alloc(n * 96)
        ctx.h2d(d_p, pts)
        check(ctx.lib.tkm_g1_bases_to_mont(ctx.h, d_p, d_p, n))
        d_s = ctx.upload_fr(frs_from_ints([k % R_MOD for k in scalars]), to_mont=False)
        t = ctypes.c_int32()
        check(ctx.lib.tkm_msm_g1_begin(ctx.h, d_s, 0, d_p, n, ctypes.byref(t)))
        ctx.dev_free(d_s)
        ctx.dev_free(d_p)
        return _PendingCommit(ctx, t.value)

    def msm_points(self, points, scalars):
        """msm_g1_bases over a handful of host points (the blinding terms of the binding)."""
        pts = np.stack([g1_from_tuple(p) for p in points])
        return g1_to_tuple(self.ctx.msm_g1_host(frs_from_ints([k % R_MOD for k in scalars]), pts))

    # ---- G1serde ops
    def g1_add(self, a, b):
        return g1_to_tuple(self.ctx.g1_add(g1_from_tuple(a), g1_from_tuple(b)))

    def g1_neg(self, a):
        from .fr import Q_MOD

        return None if a is None else (a[0], (-a[1]) % Q_MOD)

    def g1_sub(self, a, b):
        return self.g1_add(a, self.g1_neg(b))

    def g1_mul(self, a, k):
        return g1_to_tuple(self.ctx.g1_mul(g1_from_tuple(a), k % R_MOD))

    # ---- prove2's combined copy-constraint polynomial
    def p_comb(self, r, g, f, r_wX, r_wXwY, KL, K0, kappa0, x_size, y_size):
        """(r - 1) KL + kappa0 (X - 1)(r g - r(X/w,Y) f) + kappa0^2 K0 (r g - r(X/w,Y/w) f) as one fused expression in the
        evaluation domain x_size x y_size: one NTT per distinct leaf, pointwise passes, one inverse NTT
        (PolyExpr::evaluate_fused_with_domain, prove/src/lib.rs:2110-2146)."""
        from .. import PolyExpr as E

        # r(X/w, Y) and r(X/w, Y/w) are r's evaluation table rotated by x_size/m_I rows (and y_size/s_max columns): they
        # share r's leaf transform (TKM_PEX_LEAF_SHIFT), five forward biNTTs instead of seven
        m_i, s_max = x_size // 4, y_size // 2
        e_wX, e_wXwY = E.poly_over_roots(r, m_i, 0), E.poly_over_roots(r, m_i, s_max)  # == E.poly(r_wX), E.poly(r_wXwY)
        rg = E.mul(E.poly(r), E.poly(g))
        p1 = E.mul(E.sub(E.poly(r), E.scalar(1)), E.poly(KL))
        p2 = E.mul_x_minus_one(E.sub(rg, E.mul(e_wX, E.poly(f))))
        p3 = E.mul(E.poly(K0), E.sub(rg, E.mul(e_wXwY, E.poly(f))))
        expr = E.weighted_sum([(1, p1), (kappa0 % R_MOD, p2), (kappa0 * kappa0 % R_MOD, p3)])
        return expr.evaluate_fused_with_domain(x_size, y_size, self.ctx)

    # ---- the permutation polynomials s0, s1
    def permutation_polys(self, permutation, m_i, s_max, omega_m_i, omega_s_max):
        """Permutation::to_poly (libs/src/iotools/mod.rs:419-455) on the device: the identity tables w_x^row and w_y^col as
        outer products of power vectors, the listed (row, col) -> (X, Y) entries scattered over them, one inverse biNTT each.
        Nothing of size m_i * s_max is built on the host."""
        ctx, lib = self.ctx, self.ctx.lib
        n = m_i * s_max
        from .. import fr_bytes

        d_xp, d_yp = ctx.dev_alloc(m_i * 32), ctx.dev_alloc(s_max * 32)
        d_one = ctx.dev_alloc(max(m_i, s_max) * 32)
        check(lib.tkm_fr_powers(ctx.h, fr_bytes(omega_m_i)[1], d_xp, m_i))
        check(lib.tkm_fr_powers(ctx.h, fr_bytes(omega_s_max)[1], d_yp, s_max))
        check(lib.tkm_fr_vec_fill(ctx.h, fr_bytes(1)[1], d_one, max(m_i, s_max)))
        d_s0, d_s1 = ctx.dev_alloc(n * 32), ctx.dev_alloc(n * 32)
        check(lib.tkm_fr_outer_product(ctx.h, d_xp, d_one, d_s0, m_i, s_max))
        check(lib.tkm_fr_outer_product(ctx.h, d_one, d_yp, d_s1, m_i, s_max))
        if len(permutation):
            e = np.array([(p.row, p.col, p.X, p.Y) for p in permutation], dtype=np.int64)
            if e[:, 0].max() >= m_i or e[:, 2].max() >= m_i or e[:, 1].max() >= s_max or e[:, 3].max() >= s_max or e.min() < 0:
                raise ValueError("permutation entry out of range")
            idx = e[:, 0] * s_max + e[:, 1]
            _, last = np.unique(idx[::-1], return_index=True)  # a later entry for the same (row, col) overrides an earlier one
            keep = len(idx) - 1 - last
            k = len(keep)
            d_idx, d_sx, d_sy = ctx.dev_alloc(k * 4), ctx.dev_alloc(k * 4), ctx.dev_alloc(k * 4)
            ctx.h2d(d_idx, np.ascontiguousarray(idx[keep], dtype=np.uint32))
            ctx.h2d(d_sx, np.ascontiguousarray(e[keep, 2], dtype=np.uint32))
            ctx.h2d(d_sy, np.ascontiguousarray(e[keep, 3], dtype=np.uint32))
            check(lib.tkm_fr_scatter_from_table(ctx.h, d_s0, n, d_idx, d_xp, m_i, d_sx, k))
            check(lib.tkm_fr_scatter_from_table(ctx.h, d_s1, n, d_idx, d_yp, s_max, d_sy, k))
            for p_ in (d_idx, d_sx, d_sy):
                ctx.dev_free(p_)
        out = []
        for d in (d_s0, d_s1):
            h = ctypes.c_void_p()
            check(lib.tkm_poly_from_device(ctx.h, d, m_i, s_max, ctypes.byref(h)))
            q = DensePolynomialExt(ctx, h)
            check(lib.tkm_poly_ntt_inplace(ctx.h, q.h, 1, None, None))
            out.append(q)
        for p_ in (d_xp, d_yp, d_one, d_s0, d_s1):
            ctx.dev_free(p_)
        return out[0], out[1]

    # ---- prove1's recursion polynomial
    def recursion_poly(self, f, g, m_i, s_max):
        """r(X,Y) from the polynomials f, g without leaving the device (prove/src/lib.rs:1835-1880): evaluate both on the
        grid, scalers = g/f, transpose to placement-major order, exclusive suffix product, transpose back, interpolate."""
        ctx = self.ctx
        n = m_i * s_max
        if f.shape != (m_i, s_max) or g.shape != (m_i, s_max):
            return None
        fe, ge = f.clone(), g.clone()
        check(ctx.lib.tkm_poly_ntt_inplace(ctx.h, fe.h, 0, None, None))
        check(ctx.lib.tkm_poly_ntt_inplace(ctx.h, ge.h, 0, None, None))
        pf, pg = fe.device_ptr(), ge.device_ptr()
        check(ctx.lib.tkm_fr_vec_op(ctx.h, 3, pg, pf, pg, n))  # OP_DIV
        check(ctx.lib.tkm_fr_transpose(ctx.h, pg, pf, m_i, s_max))
        check(ctx.lib.tkm_fr_suffix_product(ctx.h, pf, pf, n))
        check(ctx.lib.tkm_fr_transpose(ctx.h, pf, pg, s_max, m_i))
        check(ctx.lib.tkm_poly_ntt_inplace(ctx.h, ge.h, 1, None, None))
        return ge

    def recursion_evals(self, f_evals, g_evals, m_i, s_max):
        """r(X,Y) on the grid (prove/src/lib.rs:1853-1870): scalers = g/f, transposed to placement-major order,
        r[last] = 1, r[k] = r[k+1] * scalers[k+1], transposed back.  All on the device."""
        ctx = self.ctx
        n = m_i * s_max
        d_f = ctx.upload_fr(_as_fr_array(f_evals))
        d_g = ctx.upload_fr(_as_fr_array(g_evals))
        d_t = ctx.dev_alloc(n * 32)
        check(ctx.lib.tkm_fr_vec_op(ctx.h, 3, d_g, d_f, d_g, n))  # OP_DIV
        check(ctx.lib.tkm_fr_transpose(ctx.h, d_g, d_t, m_i, s_max))
        check(ctx.lib.tkm_fr_suffix_product(ctx.h, d_t, d_f, n))
        check(ctx.lib.tkm_fr_transpose(ctx.h, d_f, d_g, s_max, m_i))
        out = ctx.download_fr(d_g, n)
        for p in (d_f, d_g, d_t):
            ctx.dev_free(p)
        return out
