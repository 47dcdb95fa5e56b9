"""Host-side mirror of the reference's `libs` API for the hot path, over the C-ABI of libtokamak_b200.

Names follow packages/backend/libs (DensePolynomialExt, BivariatePolynomial methods, Sigma1.encode_poly,
G1serde, vector_operations) so tests read like libs/src/tests.rs.  Field elements cross this layer as
Python ints or numpy uint64 arrays in the reference's canonical little-endian layout
(Fr: shape (n, 4); G1 affine: shape (n, 12), all-zero row = identity).
"""
import ctypes

import numpy as np

from . import ffi
from .ffi import TkmError, check

R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FORWARD, INVERSE = 0, 1
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 0, 1, 2, 3
PEX_LEAF, PEX_CONST, PEX_ADD, PEX_SUB, PEX_MUL, PEX_SCALE, PEX_XM1, PEX_LEAF_SHIFT = range(8)  # tkm_polyexpr_eval opcodes


def _vp(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def fr_bytes(v):
    """int | 4 x u64 array -> (keepalive numpy array, void*) of 32 canonical LE bytes, or (None, None)."""
    if v is None:
        return None, None
    if isinstance(v, (int, np.integer)):
        a = np.frombuffer(int(v % R_MOD).to_bytes(32, "little"), dtype=np.uint64).copy()
    else:
        a = np.ascontiguousarray(v, dtype=np.uint64).reshape(4)
    return a, _vp(a)


def fr_to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes(), "little")


def frs_from_ints(vs):
    return np.frombuffer(b"".join(int(v % R_MOD).to_bytes(32, "little") for v in vs), dtype=np.uint64).reshape(-1, 4).copy()


def frs_sparse(size, entries):
    """`size` field elements, zero except entries = {index: int}: the vanishing, monomial and unit-vector tables of the prover
    without a Python loop over their zeros."""
    a = np.zeros((size, 4), dtype=np.uint64)
    for i, v in entries.items():
        a[i] = np.frombuffer(int(v % R_MOD).to_bytes(32, "little"), dtype=np.uint64)
    return a


def frs_to_ints(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def _as_fr_array(a):
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return frs_from_ints(list(a))


class Context:
    """One device + one stream (utils::check_device, libs/src/utils/mod.rs:88-110). Not thread-safe."""

    def __init__(self, device=0):
        self.lib = ffi.load()
        h = ctypes.c_void_p()
        check(self.lib.tkm_ctx_create(device, ctypes.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.tkm_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- streams / timing ------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr):
        """Run the library on a caller-owned cudaStream_t.  None restores the context's own stream; the value 0
        (torch's default stream) is mapped to cudaStreamLegacy (handle 0x1) so work is ordered with it."""
        if cuda_stream_ptr is None:
            check(self.lib.tkm_ctx_set_stream(self.h, None))
        else:
            check(self.lib.tkm_ctx_set_stream(self.h, ctypes.c_void_p(cuda_stream_ptr if cuda_stream_ptr else 1)))

    def sync(self):
        check(self.lib.tkm_ctx_sync(self.h))

    def time_begin(self):
        check(self.lib.tkm_event_time_begin(self.h))

    def time_end(self):
        ms = ctypes.c_float()
        check(self.lib.tkm_event_time_end(self.h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        v = ctypes.c_uint64()
        check(self.lib.tkm_launch_count(self.h, ctypes.byref(v)))
        return v.value

    def kernel_time_last(self):
        """Milliseconds of the most recent dominant-kernel launch (k_accumulate / the k_ntt_pass launches of a biNTT)."""
        ms = ctypes.c_float()
        check(self.lib.tkm_kernel_time_last(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def poly_kernel_time_last(self):
        ms = ctypes.c_float()
        check(self.lib.tkm_poly_kernel_time_last(self.h, ctypes.byref(ms)))
        return float(ms.value)

    def msm_tree_stats(self):
        """(L, [n_0 .. n_L]): affine pair-tree levels of the last MSM accumulation pass and their entry counts."""
        lv = ctypes.c_uint32()
        cnt = np.zeros(9, dtype=np.uint64)
        check(self.lib.tkm_msm_tree_stats(self.h, ctypes.byref(lv), _vp(cnt)))
        return int(lv.value), [int(c) for c in cnt[:lv.value + 1]]

    def microbench(self, kind):
        v = ctypes.c_double()
        check(self.lib.tkm_microbench(self.h, kind, ctypes.byref(v)))
        return v.value

    # -- NTT domain (init_ntt_domain_for_size, bivariate_polynomial/mod.rs:33-55) --------------------
    def init_ntt_domain_for_size(self, size):
        if size == 0:
            raise ValueError("NTT domain size must be non-zero.")
        if size & (size - 1):
            raise ValueError("NTT domain size must be a power of two.")
        check(self.lib.tkm_ntt_domain_init(self.h, size.bit_length() - 1))

    def release_ntt_domain(self):
        check(self.lib.tkm_ntt_domain_release(self.h))

    def ntt_domain_log2(self):
        v = ctypes.c_int32()
        check(self.lib.tkm_ntt_domain_log2(self.h, ctypes.byref(v)))
        return v.value

    def get_root_of_unity(self, n):
        out = np.zeros(4, dtype=np.uint64)
        check(self.lib.tkm_root_of_unity(n.bit_length() - 1, _vp(out)))
        return fr_to_int(out)

    # -- raw device buffers -------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        check(self.lib.tkm_dev_alloc(self.h, nbytes, ctypes.byref(p)))
        return p.value

    def dev_free(self, ptr):
        check(self.lib.tkm_dev_free(self.h, ctypes.c_void_p(ptr)))

    def h2d(self, ptr, arr):
        arr = np.ascontiguousarray(arr)
        check(self.lib.tkm_memcpy_h2d(self.h, ctypes.c_void_p(ptr), _vp(arr), arr.nbytes))

    def d2h(self, arr, ptr):
        check(self.lib.tkm_memcpy_d2h(self.h, _vp(arr), ctypes.c_void_p(ptr), arr.nbytes))

    def upload_fr(self, a, to_mont=True):
        a = _as_fr_array(a)
        p = self.dev_alloc(a.nbytes)
        self.h2d(p, a)
        if to_mont:
            check(self.lib.tkm_fr_to_mont(self.h, ctypes.c_void_p(p), ctypes.c_void_p(p), a.shape[0]))
        return p

    def download_fr(self, ptr, n, from_mont=True):
        out = np.empty((n, 4), dtype=np.uint64)
        if from_mont:
            tmp = self.dev_alloc(out.nbytes)
            check(self.lib.tkm_fr_from_mont(self.h, ctypes.c_void_p(ptr), ctypes.c_void_p(tmp), n))
            self.d2h(out, tmp)
            self.dev_free(tmp)
        else:
            self.d2h(out, ptr)
        return out

    # -- host-slice forms of the ICICLE calls (HostSlice in / HostSlice out) ----------------------
    def vec_op_host(self, op, a, b):
        a, b = _as_fr_array(a), _as_fr_array(b)
        out = np.empty_like(a)
        check(self.lib.tkm_fr_vec_op_host(self.h, op, _vp(a), _vp(b), _vp(out), a.shape[0]))
        return out

    def bintt_host(self, a, x_size, y_size, direction=FORWARD, coset_x=None, coset_y=None):
        """DensePolynomialExt::_biNTT with host slices (bivariate_polynomial/mod.rs:1422-1478)."""
        a = _as_fr_array(a)
        out = np.empty_like(a)
        kx, px = fr_bytes(coset_x)
        ky, py = fr_bytes(coset_y)
        check(self.lib.tkm_bintt_host(self.h, _vp(a), _vp(out), x_size, y_size, direction, px, py))
        return out

    def bintt_dev(self, d_in, d_out, x_size, y_size, direction=FORWARD, coset_x=None, coset_y=None):
        kx, px = fr_bytes(coset_x)
        ky, py = fr_bytes(coset_y)
        check(self.lib.tkm_bintt(self.h, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), x_size, y_size, direction, px, py))

    def ntt_batch_dev(self, d_in, d_out, n, batch, columns_batch=False, direction=FORWARD, coset=None):
        k, p = fr_bytes(coset)
        check(self.lib.tkm_ntt_batch(self.h, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), n, batch, int(columns_batch), direction, p))

    def ntt_batch_scatter(self, d_in, n, batch, columns_batch, direction, coset, peer_ptrs, stride_a, stride_b, b0):
        """tkm_ntt_batch_scatter: batched 1-D transform whose last pass stores into the peers' buffers (fused exchange)."""
        k, pc = fr_bytes(coset)
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
        check(self.lib.tkm_ntt_batch_scatter(self.h, d_in, n, batch, 1 if columns_batch else 0, direction, pc, arr, len(peer_ptrs),
                                             stride_a, stride_b, b0))

    def msm_g1_host(self, scalars, bases):
        """msm::msm(HostSlice scalars, HostSlice bases, MSMConfig::default()) -> affine point (12 x u64)."""
        scalars = _as_fr_array(scalars)
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 12)
        if scalars.shape[0] != bases.shape[0]:
            raise ValueError("msm input length mismatch")
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_msm_g1_host(self.h, _vp(scalars), _vp(bases), scalars.shape[0], _vp(out)))
        return out

    def msm_g1_dev(self, d_scalars, scalars_mont, d_bases_mont, n):
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_msm_g1(self.h, ctypes.c_void_p(d_scalars), int(scalars_mont), ctypes.c_void_p(d_bases_mont), n, _vp(out)))
        return out

    def msm_g1_rect_dev(self, d_scalars, scalars_mont, s_stride, d_bases_mont, b_stride, rows, cols):
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_msm_g1_rect(self.h, ctypes.c_void_p(d_scalars), int(scalars_mont), s_stride, ctypes.c_void_p(d_bases_mont),
                                       b_stride, rows, cols, _vp(out)))
        return out

    def msm_g1_indexed_dev(self, d_scalars, scalars_mont, d_bases_mont, d_idx, n):
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_msm_g1_indexed(self.h, ctypes.c_void_p(d_scalars), int(scalars_mont), ctypes.c_void_p(d_bases_mont),
                                          ctypes.c_void_p(d_idx), n, _vp(out)))
        return out

    def transpose_dev(self, d_in, d_out, rows, cols):
        """VecOps::transpose (vector_operations/mod.rs:139,168)."""
        check(self.lib.tkm_fr_transpose(self.h, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out), rows, cols))

    def upload_bases(self, bases):
        """Canonical affine points -> device table in Montgomery form. Returns device pointer."""
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 12)
        p = self.dev_alloc(bases.nbytes)
        self.h2d(p, bases)
        check(self.lib.tkm_g1_bases_to_mont(self.h, ctypes.c_void_p(p), ctypes.c_void_p(p), bases.shape[0]))
        return p

    def g1_fixed_base_mul(self, base, scalars):
        """N scalar multiples of one base (from_coef_vec_to_g1serde_vec, iotools/mod.rs:1113-1135)."""
        base = np.ascontiguousarray(base, dtype=np.uint64).reshape(12)
        scalars = _as_fr_array(scalars)
        n = scalars.shape[0]
        ds = self.upload_fr(scalars, to_mont=False)
        dout = self.dev_alloc(n * 96)
        check(self.lib.tkm_g1_fixed_base_mul(self.h, _vp(base), ctypes.c_void_p(ds), 0, n, ctypes.c_void_p(dout)))
        out = np.empty((n, 12), dtype=np.uint64)
        self.d2h(out, dout)
        self.dev_free(ds)
        self.dev_free(dout)
        return out

    def g1_add(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12)
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(12)
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_g1_add(self.h, _vp(a), _vp(b), _vp(out)))
        return out

    def g1_sum(self, points):
        """Sum of affine points (n x 12 u64, canonical): one launch, used to combine sharded MSM partial sums."""
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_g1_sum(self.h, _vp(pts), pts.shape[0], _vp(out)))
        return out

    def g1_mul(self, a, k):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(12)
        kk, pk = fr_bytes(k)
        out = np.zeros(12, dtype=np.uint64)
        check(self.lib.tkm_g1_mul(self.h, _vp(a), pk, _vp(out)))
        return out


class Sigma1:
    """Device-resident sigma_1.xy_powers (libs/src/group_structures/mod.rs:361-394): grid rs_x x rs_y,
    index rs_y*h + i <-> x^h y^i."""

    def __init__(self, ctx, xy_powers, rs_x, rs_y):
        self.ctx = ctx
        pts = np.ascontiguousarray(xy_powers, dtype=np.uint64).reshape(-1, 12)
        assert pts.shape[0] == rs_x * rs_y
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_crs_upload(ctx.h, _vp(pts), rs_x, rs_y, ctypes.byref(h)))
        self.h, self.rs_x, self.rs_y = h, rs_x, rs_y

    def precompute(self, window_bits=20):
        """Build fixed-base tables for this CRS (for provers that reuse one CRS across many proofs)."""
        check(self.ctx.lib.tkm_crs_precompute(self.ctx.h, self.h, window_bits))

    def encode_poly(self, poly):
        """Sigma1::encode_poly (group_structures/mod.rs:59-119; iotools/mod.rs:2041-2113)."""
        out = np.zeros(12, dtype=np.uint64)
        check(self.ctx.lib.tkm_poly_commit(self.ctx.h, poly.h, self.h, _vp(out)))
        return out

    def encode_poly_begin(self, poly):
        """Queue a commitment (tkm_poly_commit_begin); returns a ticket for encode_poly_end."""
        t = ctypes.c_int32()
        check(self.ctx.lib.tkm_poly_commit_begin(self.ctx.h, poly.h, self.h, ctypes.byref(t)))
        return t.value

    def encode_poly_end(self, ticket):
        out = np.zeros(12, dtype=np.uint64)
        check(self.ctx.lib.tkm_commit_end(self.ctx.h, ticket, _vp(out)))
        return out

    def device_ptr(self):
        p = ctypes.c_void_p()
        check(self.ctx.lib.tkm_crs_device_ptr(self.h, ctypes.byref(p), None, None))
        return p.value

    def close(self):
        if self.h:
            self.ctx.lib.tkm_crs_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DensePolynomialExt:
    """Device-resident bivariate polynomial with the BivariatePolynomial trait's method names
    (libs/src/bivariate_polynomial/mod.rs:1283-1416)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_coeffs(cls, ctx, coeffs, x_size, y_size):
        a = _as_fr_array(coeffs)
        if x_size * y_size != a.shape[0]:
            raise ValueError("Mismatch between the coefficient vector and the polynomial size")
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_poly_from_coeffs_host(ctx.h, _vp(a), x_size, y_size, ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_rou_evals(cls, ctx, evals, x_size, y_size, coset_x=None, coset_y=None):
        a = _as_fr_array(evals)
        kx, px = fr_bytes(coset_x)
        ky, py = fr_bytes(coset_y)
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_poly_from_evals_host(ctx.h, _vp(a), x_size, y_size, px, py, ctypes.byref(h)))
        return cls(ctx, h)

    @classmethod
    def zero(cls, ctx, x_size=1, y_size=1):
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_poly_zero(ctx.h, x_size, y_size, ctypes.byref(h)))
        return cls(ctx, h)

    def clone(self):
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_clone(self.ctx.h, self.h, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.tkm_poly_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- shape ------------------------------------------------------------------------------------
    @property
    def shape(self):
        x, y = ctypes.c_size_t(), ctypes.c_size_t()
        check(self.ctx.lib.tkm_poly_shape(self.h, ctypes.byref(x), ctypes.byref(y)))
        return x.value, y.value

    @property
    def x_size(self):
        return self.shape[0]

    @property
    def y_size(self):
        return self.shape[1]

    def device_ptr(self):
        p = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_device_ptr(self.h, ctypes.byref(p)))
        return p.value

    def find_degree(self):
        xd, yd = ctypes.c_int64(), ctypes.c_int64()
        check(self.ctx.lib.tkm_poly_find_degree(self.ctx.h, self.h, ctypes.byref(xd), ctypes.byref(yd)))
        return xd.value, yd.value

    def is_zero(self):
        return self.find_degree() == (-1, -1)

    def resize(self, target_x_size, target_y_size):
        check(self.ctx.lib.tkm_poly_resize(self.ctx.h, self.h, target_x_size, target_y_size))

    def optimize_size(self):
        check(self.ctx.lib.tkm_poly_optimize_size(self.ctx.h, self.h))

    def mul_monomial(self, x_exponent, y_exponent):
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_mul_monomial(self.ctx.h, self.h, x_exponent, y_exponent, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    @staticmethod
    def lincomb(terms):
        """poly_comb! (prove/src/lib.rs:30-38) in one pass: sum of c * X^sx * Y^sy * p over terms (c, p) or (c, p, sx, sy)."""
        terms = [t if len(t) == 4 else (t[0], t[1], 0, 0) for t in terms]
        ctx = terms[0][1].ctx
        k = len(terms)
        hs = (ctypes.c_void_p * k)(*[t[1].h for t in terms])
        cs = frs_from_ints([int(t[0]) % R_MOD for t in terms])
        sx = np.array([t[2] for t in terms], dtype=np.uint32)
        sy = np.array([t[3] for t in terms], dtype=np.uint32)
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_poly_lincomb(ctx.h, k, hs, _vp(cs), _vp(sx), _vp(sy), ctypes.byref(h)))
        return DensePolynomialExt(ctx, h)

    # -- data movement ------------------------------------------------------------------------------
    def copy_coeffs(self):
        x, y = self.shape
        out = np.empty((x * y, 4), dtype=np.uint64)
        check(self.ctx.lib.tkm_poly_copy_coeffs_host(self.ctx.h, self.h, _vp(out)))
        return out

    def coeffs_ints(self):
        return frs_to_ints(self.copy_coeffs())

    def get_coeff(self, idx_x, idx_y):
        return self.coeffs_ints()[idx_x * self.y_size + idx_y]

    def to_rou_evals(self, coset_x=None, coset_y=None):
        x, y = self.shape
        out = np.empty((x * y, 4), dtype=np.uint64)
        kx, px = fr_bytes(coset_x)
        ky, py = fr_bytes(coset_y)
        check(self.ctx.lib.tkm_poly_to_evals_host(self.ctx.h, self.h, px, py, _vp(out)))
        return out

    # -- arithmetic -----------------------------------------------------------------------------------
    def _axpby(self, ca, other, cb):
        ka, pa = fr_bytes(ca)
        kb, pb = fr_bytes(cb)
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_axpby(self.ctx.h, self.h, pa, other.h if other is not None else None, pb, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    def __add__(self, other):
        if isinstance(other, DensePolynomialExt):
            return self._axpby(None, other, None)
        r = self.clone()
        k, p = fr_bytes(int(other))
        check(self.ctx.lib.tkm_poly_add_scalar(self.ctx.h, r.h, p))
        return r

    def __sub__(self, other):
        if isinstance(other, DensePolynomialExt):
            return self._axpby(None, other, R_MOD - 1)
        return self + ((-int(other)) % R_MOD)

    def __neg__(self):
        return self._axpby(R_MOD - 1, None, None)

    def __radd__(self, other):  # &s + &p (:1100-1140)
        return self + other

    def __rsub__(self, other):  # &s - &p: negate, then add the scalar to c00
        return (-self) + int(other)

    def __mul__(self, other):
        if isinstance(other, DensePolynomialExt):
            h = ctypes.c_void_p()
            check(self.ctx.lib.tkm_poly_mul(self.ctx.h, self.h, other.h, ctypes.byref(h)))
            return DensePolynomialExt(self.ctx, h)
        return self._axpby(int(other), None, None)

    __rmul__ = __mul__

    def scale_coeffs_x(self, s):
        return self._scale(s, None)

    def scale_coeffs_y(self, s):
        return self._scale(None, s)

    def _scale(self, sx, sy):
        kx, px = fr_bytes(sx)
        ky, py = fr_bytes(sy)
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_scale_coeffs(self.ctx.h, self.h, px, py, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    def eval(self, x, y):
        kx, px = fr_bytes(x)
        ky, py = fr_bytes(y)
        out = np.zeros(4, dtype=np.uint64)
        check(self.ctx.lib.tkm_poly_eval(self.ctx.h, self.h, px, py, _vp(out)))
        return fr_to_int(out)

    def eval_x(self, x):
        k, p = fr_bytes(x)
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_eval_x(self.ctx.h, self.h, p, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    def eval_y(self, y):
        k, p = fr_bytes(y)
        h = ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_eval_y(self.ctx.h, self.h, p, ctypes.byref(h)))
        return DensePolynomialExt(self.ctx, h)

    def div_by_vanishing_opt(self, x_degree, y_degree):
        qx, qy = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_div_by_vanishing(self.ctx.h, self.h, x_degree, y_degree, ctypes.byref(qx), ctypes.byref(qy)))
        return DensePolynomialExt(self.ctx, qx), DensePolynomialExt(self.ctx, qy)

    def div_by_vanishing(self, denom_x_degree, denom_y_degree, cache=None):
        """div_by_vanishing (:2096-2282), the legacy formulation, literally: fold the X-blocks of the numerator, divide by
        Y^d - 1 on a Y-coset of the c x (n d) grid to get Q_Y, subtract Q_Y (Y^d - 1), divide by X^c - 1 on an X-coset to
        get Q_X.  `cache` (DivByVanishingCache) keeps the coset generators and the inverted denominator tables per shape
        like the reference's.  For a numerator in the ideal it returns the polynomials of div_by_vanishing_opt (the
        decomposition with deg_X Q_Y < c is unique); the prover only calls the _opt form."""
        c, d = int(denom_x_degree), int(denom_y_degree)
        if c <= 0 or d <= 0 or c & (c - 1) or d & (d - 1):
            raise ValueError("The denominators must have degress as powers of two.")
        self.optimize_size()
        xs, ys = self.shape
        xd, yd = self.find_degree()
        if xd < c or yd < d:
            raise ValueError("The numerator must have grater degrees than denominators.")
        m, n = xs // c, ys // d
        cache = cache if cache is not None else DivByVanishingCache()
        ctx = self.ctx
        hit_x = cache.find(cache.denom_x_eval_inv, m * c, n * d, c)
        zeta = hit_x["coset"] if hit_x else cache.fresh_coset()
        hit_y = cache.find(cache.denom_y_eval_inv, c, n * d, d)
        xi = hit_y["coset"] if hit_y else zeta

        def build_denom_inv(target_x, target_y, base, coset, along_y):
            axis = target_y if along_y else target_x
            root = ctx.get_root_of_unity(axis // base)
            cp = pow(coset, base, R_MOD)
            vals, w = [], 1
            for _ in range(axis):
                vals.append((cp * w - 1) % R_MOD)
                w = w * root % R_MOD
            ax = frs_from_ints(vals)
            mat = np.tile(ax, (target_x, 1)) if along_y else np.repeat(ax, target_y, axis=0)
            return ctx.vec_op_host(OP_DIV, frs_from_ints([1] * (target_x * target_y)), mat)

        blocks = self.copy_coeffs().reshape(m, c * ys, 4)
        acc = np.ascontiguousarray(blocks[0])
        for i in range(1, m):
            acc = ctx.vec_op_host(OP_ADD, acc, np.ascontiguousarray(blocks[i]))
        r_tilde = DensePolynomialExt.from_coeffs(ctx, acc, c, n * d).to_rou_evals(None, xi)
        if not hit_y:
            hit_y = {"coset": xi, "x_size": c, "y_size": n * d, "base": d, "evals": build_denom_inv(c, n * d, d, xi, True)}
            cache.denom_y_eval_inv.append(hit_y)
        quo_y = DensePolynomialExt.from_rou_evals(ctx, ctx.vec_op_host(OP_MUL, r_tilde, hit_y["evals"]), c, n * d, None, xi)
        r = quo_y.mul_monomial(0, d) - quo_y
        b = self - r
        b.resize(m * c, n * d)
        b_tilde = b.to_rou_evals(zeta, None)
        if not hit_x:
            hit_x = {"coset": zeta, "x_size": m * c, "y_size": n * d, "base": c, "evals": build_denom_inv(m * c, n * d, c, zeta, False)}
            cache.denom_x_eval_inv.append(hit_x)
        quo_x = DensePolynomialExt.from_rou_evals(ctx, ctx.vec_op_host(OP_MUL, b_tilde, hit_x["evals"]), m * c, n * d, zeta, None)
        return quo_x, quo_y

    def _divide_uni(self, denominator, y_dir):
        q, r = ctypes.c_void_p(), ctypes.c_void_p()
        check(self.ctx.lib.tkm_poly_divide_uni(self.ctx.h, self.h, denominator.h, 1 if y_dir else 0, ctypes.byref(q), ctypes.byref(r)))
        return DensePolynomialExt(self.ctx, q), DensePolynomialExt(self.ctx, r)

    def divide_x(self, denominator):
        """divide_x (:1998-2023): long division of every Y-column along X by an X-univariate denominator -> (quotient, remainder)."""
        return self._divide_uni(denominator, False)

    def divide_y(self, denominator):
        """divide_y (:2025-2050)."""
        return self._divide_uni(denominator, True)

    def get_univariate_polynomial_x(self, idx_y):
        """The X-univariate polynomial of the idx_y-th power of Y (:1760-1770), shape x_size x 1."""
        x, y = self.shape
        return DensePolynomialExt.from_coeffs(self.ctx, np.ascontiguousarray(self.copy_coeffs().reshape(x, y, 4)[:, idx_y]), x, 1)

    def get_univariate_polynomial_y(self, idx_x):
        """The Y-univariate polynomial of the idx_x-th power of X (:1772-1782), shape 1 x y_size."""
        x, y = self.shape
        return DensePolynomialExt.from_coeffs(self.ctx, np.ascontiguousarray(self.copy_coeffs().reshape(x, y, 4)[idx_x]), 1, y)

    def degree(self):
        return self.find_degree()

    def div_by_ruffini(self, x, y):
        kx, px = fr_bytes(x)
        ky, py = fr_bytes(y)
        qx, qy = ctypes.c_void_p(), ctypes.c_void_p()
        r = np.zeros(4, dtype=np.uint64)
        check(self.ctx.lib.tkm_poly_div_by_ruffini(self.ctx.h, self.h, px, py, ctypes.byref(qx), ctypes.byref(qy), _vp(r)))
        return DensePolynomialExt(self.ctx, qx), DensePolynomialExt(self.ctx, qy), fr_to_int(r)


class DivByVanishingCache:
    """DivByVanishingCache / DenomCache (bivariate_polynomial/mod.rs:88-110): inverted denominator evaluations and the coset
    generator they were built for, keyed by (x_size, y_size, base)."""

    def __init__(self, seed=None):
        self.denom_x_eval_inv, self.denom_y_eval_inv = [], []
        self._rng = __import__("random").Random(seed)

    @staticmethod
    def find(entries, x_size, y_size, base):
        for e in entries:
            if (e["x_size"], e["y_size"], e["base"]) == (x_size, y_size, base):
                return e
        return None

    def fresh_coset(self):
        return self._rng.randrange(2, R_MOD)


def _domain_size_for_degree(degree):
    """domain_size_for_degree (bivariate_polynomial/mod.rs:438-444)."""
    if degree < 0:
        return 1
    n = degree + 1
    return 1 << (n - 1).bit_length()


class PolyExpr:
    """Expression DAG over polynomials evaluated either coefficient-wise or fused in the evaluation domain:
    one NTT per distinct leaf (cached by object identity), pointwise device ops, one inverse NTT at the end
    (PolyExpr, libs/src/bivariate_polynomial/mod.rs:140-436; used for prove2's p_comb, prove/src/lib.rs:2110-2146)."""

    def __init__(self, kind, *args):
        self.kind, self.args = kind, args

    # -- constructors (same names as the reference) -------------------------------------------------
    @staticmethod
    def poly(p):
        return PolyExpr("poly", p)

    @staticmethod
    def poly_over_roots(p, mx=0, my=0):
        """p(X / w_mx, Y / w_my) with w_m the primitive m-th root of unity (m a power of two, 0 = axis not scaled): what
        scale_coeffs_x / _y by an inverse root produce (r(X/w, Y), r(X/w, Y/w) in prove2, prove/src/lib.rs:2110-2146), as a leaf
        that shares p's transform -- on the evaluation grid it is p's table rotated (TKM_PEX_LEAF_SHIFT).  An extension of the
        reference's PolyExpr::poly; evaluate_coeffs falls back to the scaled polynomial."""
        for m in (mx, my):
            if m and m & (m - 1):
                raise ValueError("the root's order must be a power of two")
        return PolyExpr("poly_shift", p, int(mx), int(my))

    @staticmethod
    def scalar(s):
        return PolyExpr("scalar", int(s) % R_MOD)

    @staticmethod
    def add(lhs, rhs):
        return PolyExpr("add", lhs, rhs)

    @staticmethod
    def sub(lhs, rhs):
        return PolyExpr("sub", lhs, rhs)

    @staticmethod
    def mul(lhs, rhs):
        return PolyExpr("mul", lhs, rhs)

    @staticmethod
    def scale(s, expr):
        return PolyExpr("scale", int(s) % R_MOD, expr)

    @staticmethod
    def mul_x_minus_one(expr):
        return PolyExpr("xm1", expr)

    @staticmethod
    def weighted_sum(terms):
        return PolyExpr("sum", [PolyExpr.scale(s, e) for s, e in terms])

    def _ctx(self):
        if self.kind in ("poly", "poly_shift"):
            return self.args[0].ctx
        for a in self.args:
            if isinstance(a, PolyExpr):
                c = a._ctx()
                if c is not None:
                    return c
            if isinstance(a, list):
                for t in a:
                    c = t._ctx()
                    if c is not None:
                        return c
        return None

    # -- coefficient-domain evaluation (evaluate_coeffs, :190-218) --------------------------------------
    def evaluate_coeffs(self, ctx=None):
        ctx = ctx or self._ctx()
        k, a = self.kind, self.args
        if k == "poly":
            return a[0].clone()
        if k == "poly_shift":
            q = a[0].clone()
            if a[1]:
                q = q.scale_coeffs_x(pow(ctx.get_root_of_unity(a[1]), R_MOD - 2, R_MOD))
            if a[2]:
                q = q.scale_coeffs_y(pow(ctx.get_root_of_unity(a[2]), R_MOD - 2, R_MOD))
            return q
        if k == "scalar":
            return DensePolynomialExt.from_coeffs(ctx, frs_from_ints([a[0]]), 1, 1)
        if k == "add":
            return a[0].evaluate_coeffs(ctx) + a[1].evaluate_coeffs(ctx)
        if k == "sub":
            return a[0].evaluate_coeffs(ctx) - a[1].evaluate_coeffs(ctx)
        if k == "mul":
            return a[0].evaluate_coeffs(ctx) * a[1].evaluate_coeffs(ctx)
        if k == "scale":
            return a[1].evaluate_coeffs(ctx) * a[0]
        if k == "xm1":
            p = a[0].evaluate_coeffs(ctx)
            return p.mul_monomial(1, 0) - p
        if k == "sum":
            if not a[0]:
                return DensePolynomialExt.zero(ctx)
            acc = a[0][0].evaluate_coeffs(ctx)
            for t in a[0][1:]:
                acc = acc + t.evaluate_coeffs(ctx)
            return acc
        raise ValueError(k)

    # -- degree bound (:262-309) ---------------------------------------------------------------------------
    def degree_bound(self):
        k, a = self.kind, self.args
        if k in ("poly", "poly_shift"):
            return a[0].find_degree()
        if k == "scalar":
            return (-1, -1) if a[0] == 0 else (0, 0)
        if k in ("add", "sub"):
            l, r = a[0].degree_bound(), a[1].degree_bound()
            return max(l[0], r[0]), max(l[1], r[1])
        if k == "mul":
            l, r = a[0].degree_bound(), a[1].degree_bound()
            if min(l + r) < 0:
                return -1, -1
            return l[0] + r[0], l[1] + r[1]
        if k == "scale":
            return (-1, -1) if a[0] == 0 else a[1].degree_bound()
        if k == "xm1":
            d = a[0].degree_bound()
            return (-1, -1) if min(d) < 0 else (d[0] + 1, d[1])
        if k == "sum":
            out = (-1, -1)
            for t in a[0]:
                d = t.degree_bound()
                out = (max(out[0], d[0]), max(out[1], d[1]))
            return out
        raise ValueError(k)

    # -- fused evaluation (:220-260, 311-435) ------------------------------------------------------------------
    def evaluate_fused(self, ctx=None):
        xd, yd = self.degree_bound()
        return self.evaluate_fused_with_domain(_domain_size_for_degree(xd), _domain_size_for_degree(yd), ctx)

    def evaluate_fused_with_domain(self, target_x_size, target_y_size, ctx=None):
        ctx = ctx or self._ctx()
        if ctx is None:
            raise ValueError("an expression without polynomial leaves needs an explicit context")
        if target_x_size & (target_x_size - 1) or target_y_size & (target_y_size - 1):
            raise ValueError("Fused polynomial expression domains must be powers of two.")
        xd, yd = self.degree_bound()
        if _domain_size_for_degree(xd) > target_x_size or _domain_size_for_degree(yd) > target_y_size:
            raise ValueError("Fused polynomial expression domain is too small for the expression degree.")
        leaves, consts, prog = [], [], []

        def leaf(p):
            for i, q in enumerate(leaves):  # pointer-keyed leaf cache (:459-502): one forward NTT per distinct polynomial
                if q is p:
                    return i
            leaves.append(p)
            return len(leaves) - 1

        def const(v):
            if v not in consts:
                consts.append(v)
            return consts.index(v)

        def emit(e):
            k, a = e.kind, e.args
            if k == "poly":
                prog.append(PEX_LEAF | leaf(a[0]) << 8)
            elif k == "poly_shift":  # p(X / w_mx, Y / w_my): the same leaf transform, read rotated
                if a[1] > target_x_size or a[2] > target_y_size:
                    raise ValueError("the root's order must divide the domain's extent")
                fx = (a[1].bit_length() if a[1] else 0)  # log2(m) + 1 for a power of two, 0 = axis not scaled
                fy = (a[2].bit_length() if a[2] else 0)
                prog.append(PEX_LEAF_SHIFT | (leaf(a[0]) | fx << 4 | fy << 10) << 8)
            elif k == "scalar":
                prog.append(PEX_CONST | const(a[0]) << 8)
            elif k in ("add", "sub", "mul"):
                emit(a[0])
                emit(a[1])
                prog.append({"add": PEX_ADD, "sub": PEX_SUB, "mul": PEX_MUL}[k])
            elif k == "scale":
                emit(a[1])
                if a[0] != 1:
                    prog.append(PEX_SCALE | const(a[0]) << 8)
            elif k == "xm1":
                emit(a[0])
                prog.append(PEX_XM1)
            elif k == "sum":
                if not a[0]:
                    prog.append(PEX_CONST | const(0) << 8)
                for n_, t in enumerate(a[0]):
                    emit(t)
                    if n_:
                        prog.append(PEX_ADD)
            else:
                raise ValueError(k)

        emit(self)
        hs = (ctypes.c_void_p * max(1, len(leaves)))(*[q.h for q in leaves])
        cs = frs_from_ints(consts or [0])
        pr = np.array(prog, dtype=np.uint32)
        h = ctypes.c_void_p()
        check(ctx.lib.tkm_polyexpr_eval(ctx.h, hs, len(leaves), _vp(pr), len(prog), _vp(cs), len(consts), target_x_size, target_y_size, ctypes.byref(h)))
        return DensePolynomialExt(ctx, h)
