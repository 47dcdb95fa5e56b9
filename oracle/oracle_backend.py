"""CPU twin of tokamak_b200.protocol.backend.GpuBackend over the oracle (TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs).

It implements the same small backend interface with oracle/oracle.c (OpenMP) and Python integers, so the protocol
driver can be run end to end on the host: the proofs of the two backends must be byte-identical, and the restated
verifier must accept both.  Polynomials are (x*y, 4) uint64 arrays of canonical little-endian limbs, row-major with X
as the row index (libs/src/bivariate_polynomial/mod.rs:1756)."""
import numpy as np

import oracle_ffi as O
import pyref as P

R_MOD = P.R_MOD


def _fr(v):
    return O.fr_from_int(v % R_MOD)


def _pow2(v):
    r = 1
    while r < v:
        r <<= 1
    return r


def _pad(a, x, y, tx, ty):
    """resize (bivariate_polynomial/mod.rs:1784-1806): exact crop / zero-pad."""
    if (x, y) == (tx, ty):
        return a
    out = np.zeros((tx, ty, 4), dtype=np.uint64)
    cx, cy = min(x, tx), min(y, ty)
    out[:cx, :cy] = a.reshape(x, y, 4)[:cx, :cy]
    return out.reshape(tx * ty, 4)


def uvw_evals(params, placements, r1cs_list):
    """Evaluation tables of u, v, w on the n x s_max grid, row-major [row][placement] (read_R1CS_gen_uvwXY +
    eval_uvwxy_sparse_rows, iotools/mod.rs:1287-1420; the transpose at :1363-1365 is folded into the indexing).
    Returns three (n*s_max, 4) uint64 arrays of canonical little-endian limbs."""
    n, s_max = params.n, params.s_max
    if len(placements) > s_max:
        raise ValueError("placement_variables length exceeds s_max.")
    out = [np.zeros((n * s_max, 4), dtype=np.uint64) for _ in range(3)]
    mask = (1 << 64) - 1
    for col, pl in enumerate(placements):
        var = pl.variables
        for row, abc in enumerate(r1cs_list[pl.subcircuitId].constraints):
            for m in range(3):
                lc = abc[m]
                if not lc:
                    continue
                acc = 0
                for wire, coeff in lc:
                    acc += coeff * var[wire]
                acc %= R_MOD
                if acc:
                    out[m][row * s_max + col] = (acc & mask, (acc >> 64) & mask, (acc >> 128) & mask, acc >> 192)
    return out


class OraclePoly:
    def __init__(self, data, x, y):
        self.a = np.ascontiguousarray(data, dtype=np.uint64).reshape(x * y, 4)
        self.x, self.y = x, y

    @property
    def shape(self):
        return self.x, self.y

    def clone(self):
        return OraclePoly(self.a.copy(), self.x, self.y)

    def find_degree(self):
        nz = self.a.reshape(self.x, self.y, 4).any(axis=2)
        if not nz.any():
            return -1, -1
        return int(np.nonzero(nz.any(axis=1))[0].max()), int(np.nonzero(nz.any(axis=0))[0].max())

    def copy_coeffs(self):
        return self.a.copy()

    def _binary(self, other, op):
        tx, ty = max(self.x, other.x), max(self.y, other.y)
        return OraclePoly(O.fr_vec_op(op, _pad(self.a, self.x, self.y, tx, ty), _pad(other.a, other.x, other.y, tx, ty)), tx, ty)

    def __add__(self, other):
        if isinstance(other, OraclePoly):
            return self._binary(other, "add")
        r = self.clone()
        r.a[0] = _fr(O.fr_to_int(r.a[0]) + int(other))
        return r

    def __sub__(self, other):
        if isinstance(other, OraclePoly):
            return self._binary(other, "sub")
        return self + ((-int(other)) % R_MOD)

    def __neg__(self):
        return self * (R_MOD - 1)

    def __mul__(self, other):
        if isinstance(other, OraclePoly):
            tx, ty = _pow2(self.x + other.x - 1), _pow2(self.y + other.y - 1)
            return OraclePoly(O.poly_mul_padded(_pad(self.a, self.x, self.y, tx, ty), _pad(other.a, other.x, other.y, tx, ty), tx, ty), tx, ty)
        k = np.tile(_fr(int(other)), (self.x * self.y, 1))
        return OraclePoly(O.fr_vec_op("mul", self.a, k), self.x, self.y)

    __rmul__ = __mul__

    def mul_monomial(self, ex, ey):
        tx, ty = _pow2(self.x + ex), _pow2(self.y + ey)
        out = np.zeros((tx, ty, 4), dtype=np.uint64)
        out[ex:ex + self.x, ey:ey + self.y] = self.a.reshape(self.x, self.y, 4)
        return OraclePoly(out.reshape(tx * ty, 4), tx, ty)

    def scale_coeffs_x(self, s):
        return OraclePoly(O.scale_coeffs(self.a, self.x, self.y, sx=_fr(s)), self.x, self.y)

    def scale_coeffs_y(self, s):
        return OraclePoly(O.scale_coeffs(self.a, self.x, self.y, sy=_fr(s)), self.x, self.y)

    def eval(self, px, py):
        return O.fr_to_int(O.eval_xy(self.a, self.x, self.y, _fr(px), _fr(py)))

    def to_rou_evals(self):
        return O.bintt(self.a, self.x, self.y, False)

    def div_by_vanishing_opt(self, c, d):
        qx, qy = O.div_by_vanishing_opt(self.a, self.x, self.y, c, d)
        return OraclePoly(qx, self.x, self.y), OraclePoly(qy, c, self.y)

    def div_by_ruffini(self, px, py):
        qx, qy, r = O.div_by_ruffini(self.a, self.x, self.y, _fr(px), _fr(py))
        return OraclePoly(qx, self.x, self.y), OraclePoly(qy, 1, self.y), O.fr_to_int(r)


class OracleTable:
    def __init__(self, points, rows, cols):
        self.points = np.ascontiguousarray(points, dtype=np.uint64).reshape(rows * cols, 12)
        self.rows, self.cols = rows, cols

    def points_host(self):
        return self.points


class OracleBackend:
    name = "oracle"

    def __init__(self):
        O.build()

    def init_ntt_domain(self, size):
        pass

    def uvw_polys(self, params, csr, wt):
        """read_R1CS_gen_uvwXY literally: per-placement sparse-row dot products on Python integers, then three INTTs."""
        u, v, w = uvw_evals(params, wt.placements, csr.r1cs_list)
        return tuple(self.from_rou_evals(e, params.n, params.s_max) for e in (u, v, w))

    def from_coeffs(self, coeffs, x, y):
        return OraclePoly(np.array(coeffs, dtype=np.uint64, copy=True), x, y)

    def from_rou_evals(self, evals, x, y):
        return OraclePoly(O.bintt(np.ascontiguousarray(evals, dtype=np.uint64), x, y, True), x, y)

    def make_table(self, col, row, base):
        c = np.repeat(O.frs_from_ints([v % R_MOD for v in col]), len(row), axis=0)
        r = np.tile(O.frs_from_ints([v % R_MOD for v in row]), (len(col), 1))
        return OracleTable(O.g1_fixed_base_mul_batch(O.g1_from_tuple(base), O.fr_vec_op("mul", c, r)), len(col), len(row))

    def table_from_points(self, points, rows, cols):
        return OracleTable(points, rows, cols)

    def commit(self, table, poly):
        """Sigma1::encode_poly: optimize_size, CRS bound check, MSM over the trimmed rectangle
        (iotools/mod.rs:2041-2113)."""
        dx, dy = poly.find_degree()
        if dx < 0:
            return None
        rows, cols = _pow2(dx + 1), _pow2(dy + 1)
        if rows > table.rows or cols > table.cols:
            raise ValueError("Insufficient length of sigma.sigma_1.xy_powers")
        a = _pad(poly.a, poly.x, poly.y, max(rows, poly.x), max(cols, poly.y))
        return O.g1_to_tuple(O.msm_g1_rect(a, max(cols, poly.y), table.points, table.cols, rows, cols))

    def msm_indexed(self, table, idx, scalars):
        idx = np.asarray(idx, dtype=np.int64)
        if idx.shape[0] == 0:
            return None
        return O.g1_to_tuple(O.msm_g1(np.ascontiguousarray(scalars, dtype=np.uint64), np.ascontiguousarray(table.points[idx])))

    def msm_points(self, points, scalars):
        pts = np.stack([O.g1_from_tuple(p) for p in points])
        return O.g1_to_tuple(O.msm_g1(O.frs_from_ints([k % R_MOD for k in scalars]), pts))

    def g1_add(self, a, b):
        return O.g1_to_tuple(O.g1_add(O.g1_from_tuple(a), O.g1_from_tuple(b)))

    def g1_neg(self, a):
        return P.g1_neg(a)

    def g1_sub(self, a, b):
        return self.g1_add(a, self.g1_neg(b))

    def g1_mul(self, a, k):
        return O.g1_to_tuple(O.g1_mul(O.g1_from_tuple(a), _fr(k)))

    def recursion_evals(self, f_evals, g_evals, m_i, s_max):
        """prove/src/lib.rs:1853-1870, literally: div, transpose, serial suffix product, transpose back."""
        sc = O.fr_vec_op("mul", np.ascontiguousarray(g_evals, dtype=np.uint64), O.fr_vec_inv(np.ascontiguousarray(f_evals, dtype=np.uint64)))
        sc_tr = O.frs_to_ints(np.ascontiguousarray(sc.reshape(m_i, s_max, 4).transpose(1, 0, 2)).reshape(m_i * s_max, 4))
        n = m_i * s_max
        r = [0] * n
        r[n - 1] = 1
        for i in range(n - 2, -1, -1):
            r[i] = r[i + 1] * sc_tr[i + 1] % R_MOD
        r_tr = O.frs_from_ints(r).reshape(s_max, m_i, 4).transpose(1, 0, 2)
        return np.ascontiguousarray(r_tr).reshape(n, 4)
