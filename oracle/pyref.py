"""Python big-integer restatement of the Tokamak zk-EVM hot path (TEST INFRASTRUCTURE ONLY).

This file is part of the *oracle*: it is imported only by tests/, by
__graft_entry__.smoke() and by the golden-vector generator.  Nothing in the
product path (tokamak-zk-evm_b200/) may import it.

PARITY UNPINNED: the reference (tokamak-network/Tokamak-zk-EVM) ships no golden
vectors / known-answer tests for NTT, MSM or proofs; its
tests are algebraic identities on random inputs, and the arithmetic lives in
the un-vendored ICICLE v3.8.0 crates (packages/backend/Cargo.toml:20-23).  The
functions below restate the reference's algorithms from its own call sites and
the identities its tests pin; the one free convention (the primitive 2^32-th
root of unity) is the 5-based root (see ROU below).

Every function cites the reference file:line it follows (paths relative to
/root/reference/packages/backend/).  Pure-Python loops: use for small cases.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------
# Fields.  BLS12-381: Fr (scalar field, 255 bit, 2-adicity 32), Fq (base field).
# ----------------------------------------------------------------------------
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
TWO_ADICITY = 32
# ICICLE derives omega_{2^k} = rou^(2^(32-k)) from one fixed 2^32-th root.  In-tree
# evidence for the 5-based root: the browser prover takes roots from ffjavascript
# Fr.w[] (backend-wasm/src/runtime/field/field-runtime.ts:118-124), built from the
# smallest quadratic non-residue (5), and must verify native proofs.
ROU = pow(5, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
assert ROU == 0x0212D79E5B416B6F0FD56DC8D168D6C0C4024FF270B3E0941B788F500B912F1F

# Standard BLS12-381 G1 generator and the reference's --fixed-tau generator
# (setup/trusted-setup/src/main.rs:71-74).
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
G1_GEN_FIXED_TAU = (
    0x0B001B4CC05FA01578BE7D4E821D6FF58F2A05C584FBA3CB31A37942DECE65EADEC9A878ADD2282F7C2513ABB8D4AB05,
    0x15E237775397ED22EEF43DD36CDCA277C9CF6FA7E4FFFF0A5BB4B20A82392CAACF0F63FB6CDB02BCCF2F5AF14970D6B9,
)
# libs/src/field_structures/mod.rs:43-65 (Tau::gen_fixed)
TAU_FIXED = {
    "x": 0x7234CD9B97845E0125E84AE3AE81354E004558D8C82A83425652BC7B9ED49F7D,
    "y": 0x6ED0EEA55CBEEEBDC7A41033EBD196FFECC1806FDBC13A8D41B8F1AA273A4037,
    "alpha": 0x7234CD9B97845E0125E84AE3AE81354E004558D8C82A83425652BC7B9ED49F7D,
    "gamma": 0x088DFE3D1B76775EC267D6D0E27B753EC904C76E0BC32CA8223DC2AE1A0AC6B4,
    "delta": 0x04B8CE26374C547D8722AC51F5ED1E0F9CB891C332C69C865D96AF150189A818,
    "eta": 0x52EB2AEB35B72B94A19EA232E984850F2CDA5542FDC10368955D8AC6274F8579,
}


def fr_inv(a: int) -> int:
    """ICICLE convention inv(0) = 0 (libs/src/bivariate_polynomial/mod.rs:2011-2013)."""
    return pow(a, R_MOD - 2, R_MOD)


def root_of_unity(n: int) -> int:
    """ntt::get_root_of_unity(n) for power-of-two n (bivariate_polynomial/mod.rs:50,505)."""
    assert n > 0 and n & (n - 1) == 0 and n <= (1 << TWO_ADICITY)
    log_n = n.bit_length() - 1
    return pow(ROU, 1 << (TWO_ADICITY - log_n), R_MOD)


# ----------------------------------------------------------------------------
# Byte formats at the boundary: 32-byte LE canonical Fr,
# 96-byte (x || y) LE canonical affine G1 with all-zero = identity.
# ----------------------------------------------------------------------------
def fr_to_bytes(a: int) -> bytes:
    return (a % R_MOD).to_bytes(32, "little")


def fr_from_bytes(b: bytes) -> int:
    return int.from_bytes(b, "little")


def frs_to_bytes(v) -> bytes:
    return b"".join(fr_to_bytes(a) for a in v)


def frs_from_bytes(b: bytes):
    return [int.from_bytes(b[i : i + 32], "little") for i in range(0, len(b), 32)]


def g1_to_bytes(pt) -> bytes:
    if pt is None:
        return bytes(96)
    return pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def g1_from_bytes(b: bytes):
    x = int.from_bytes(b[:48], "little")
    y = int.from_bytes(b[48:96], "little")
    if x == 0 and y == 0:
        return None
    return (x, y)


# ----------------------------------------------------------------------------
# 1-D / 2-D NTT with ICICLE semantics:
# natural order in and out; inverse includes 1/N; coset forward = scale coeff i
# by g^i then NTT; coset inverse = INTT then scale by g^-i  (libs/src/tests.rs:134-180).
# ----------------------------------------------------------------------------
def _ntt_inplace(a, omega):
    n = len(a)
    j = 0
    for i in range(1, n):
        bit = n >> 1
        while j & bit:
            j ^= bit
            bit >>= 1
        j ^= bit
        if i < j:
            a[i], a[j] = a[j], a[i]
    length = 2
    while length <= n:
        w_len = pow(omega, n // length, R_MOD)
        half = length >> 1
        for start in range(0, n, length):
            w = 1
            for k in range(start, start + half):
                u = a[k]
                v = a[k + half] * w % R_MOD
                a[k] = (u + v) % R_MOD
                a[k + half] = (u - v) % R_MOD
                w = w * w_len % R_MOD
        length <<= 1


def ntt(vals, inverse=False, coset=None):
    """ICICLE ntt::ntt for one vector (call sites bivariate_polynomial/mod.rs:1449-1477)."""
    n = len(vals)
    a = [v % R_MOD for v in vals]
    if n == 1:
        return a
    omega = root_of_unity(n)
    if not inverse:
        if coset is not None and coset != 1:
            g = 1
            for i in range(n):
                a[i] = a[i] * g % R_MOD
                g = g * coset % R_MOD
        _ntt_inplace(a, omega)
    else:
        _ntt_inplace(a, fr_inv(omega))
        n_inv = fr_inv(n)
        a = [x * n_inv % R_MOD for x in a]
        if coset is not None and coset != 1:
            gi = fr_inv(coset)
            g = 1
            for i in range(n):
                a[i] = a[i] * g % R_MOD
                g = g * gi % R_MOD
    return a


def bintt(mat, x_size, y_size, inverse=False, coset_x=None, coset_y=None):
    """DensePolynomialExt::_biNTT (bivariate_polynomial/mod.rs:1422-1478).

    Row-major, X = row index, Y = contiguous column index.  Y pass first (row
    batch), then X pass (columns_batch).  Degenerate axes take the 1-D path.
    """
    assert len(mat) == x_size * y_size
    if x_size == 1:
        return ntt(mat, inverse, coset_y)
    if y_size == 1:
        return ntt(mat, inverse, coset_x)
    out = [0] * (x_size * y_size)
    for i in range(x_size):
        out[i * y_size : (i + 1) * y_size] = ntt(mat[i * y_size : (i + 1) * y_size], inverse, coset_y)
    for j in range(y_size):
        col = ntt(out[j::y_size], inverse, coset_x)
        out[j::y_size] = col
    return out


# ----------------------------------------------------------------------------
# Bivariate polynomial helpers (coefficients row-major [x][y]).
# ----------------------------------------------------------------------------
def next_pow2_size(t: int) -> int:
    """_find_size_as_twopower (bivariate_polynomial/mod.rs:72-86)."""
    assert t > 0
    if t & (t - 1) == 0:
        return t
    return 1 << t.bit_length()


def find_degree(c, x_size, y_size):
    """find_degree (bivariate_polynomial/mod.rs:1480-1515): (-1,-1) for the zero polynomial."""
    xd = -1
    for i in range(x_size - 1, -1, -1):
        if any(c[i * y_size + j] % R_MOD for j in range(y_size)):
            xd = i
            break
    yd = -1
    for j in range(y_size - 1, -1, -1):
        if any(c[i * y_size + j] % R_MOD for i in range(x_size)):
            yd = j
            break
    return xd, yd


def resize(c, x_size, y_size, tx, ty):
    """resize (bivariate_polynomial/mod.rs:1784-1806): crop/zero-pad to next pow2 of (tx,ty)."""
    nx, ny = next_pow2_size(tx), next_pow2_size(ty)
    out = [0] * (nx * ny)
    for i in range(min(x_size, nx)):
        w = min(y_size, ny)
        out[i * ny : i * ny + w] = c[i * y_size : i * y_size + w]
    return out, nx, ny


def resize_exact(c, x_size, y_size, tx, ty):
    """vector_operations::resize (vector_operations/mod.rs:653-672): exact (tx,ty) rectangle."""
    out = [0] * (tx * ty)
    for i in range(min(x_size, tx)):
        w = min(y_size, ty)
        out[i * ty : i * ty + w] = c[i * y_size : i * y_size + w]
    return out


def mul_monomial(c, x_size, y_size, ex, ey):
    """mul_monomial (bivariate_polynomial/mod.rs:1820-1844) with degree = size-1."""
    nx, ny = next_pow2_size(x_size + ex), next_pow2_size(y_size + ey)
    out = [0] * (nx * ny)
    for i in range(x_size):
        out[ny * (i + ex) + ey : ny * (i + ex) + ey + y_size] = c[i * y_size : (i + 1) * y_size]
    return out, nx, ny


def scale_coeffs(c, x_size, y_size, sx=None, sy=None):
    """scale_coeffs_x / scale_coeffs_y (bivariate_polynomial/mod.rs:1553-1613): c_ij * sx^i * sy^j."""
    out = list(c)
    if sx is not None:
        f = 1
        for i in range(x_size):
            for j in range(y_size):
                out[i * y_size + j] = out[i * y_size + j] * f % R_MOD
            f = f * sx % R_MOD
    if sy is not None:
        pw = [1] * y_size
        for j in range(1, y_size):
            pw[j] = pw[j - 1] * sy % R_MOD
        for i in range(x_size):
            for j in range(y_size):
                out[i * y_size + j] = out[i * y_size + j] * pw[j] % R_MOD
    return out


def eval_xy(c, x_size, y_size, x, y):
    """eval (bivariate_polynomial/mod.rs:1719-1750): P(x, y)."""
    acc = 0
    for i in range(x_size - 1, -1, -1):
        row = 0
        for j in range(y_size - 1, -1, -1):
            row = (row * y + c[i * y_size + j]) % R_MOD
        acc = (acc * x + row) % R_MOD
    return acc


def eval_x(c, x_size, y_size, x):
    """eval_x (bivariate_polynomial/mod.rs:1719-1729): returns the Y-polynomial P(x, Y), length y_size."""
    out = [0] * y_size
    for j in range(y_size):
        acc = 0
        for i in range(x_size - 1, -1, -1):
            acc = (acc * x + c[i * y_size + j]) % R_MOD
        out[j] = acc
    return out


def eval_y(c, x_size, y_size, y):
    """eval_y (bivariate_polynomial/mod.rs:1731-1740): returns the X-polynomial P(X, y), length x_size."""
    out = [0] * x_size
    for i in range(x_size):
        acc = 0
        for j in range(y_size - 1, -1, -1):
            acc = (acc * y + c[i * y_size + j]) % R_MOD
        out[i] = acc
    return out


def poly_mul(a, ax, ay, b, bx, by):
    """_mul (bivariate_polynomial/mod.rs:1846-1996) for the generic (non-constant) case.

    Returns (coeffs, x_size, y_size) with the padded power-of-two shape the reference keeps.
    """
    adx, ady = find_degree(a, ax, ay)
    bdx, bdy = find_degree(b, bx, by)
    tx, ty = adx + bdx + 1, ady + bdy + 1
    ea, nx, ny = resize(a, ax, ay, tx, ty)
    eb, _, _ = resize(b, bx, by, tx, ty)
    fa = bintt(ea, nx, ny)
    fb = bintt(eb, nx, ny)
    prod = [u * v % R_MOD for u, v in zip(fa, fb)]
    return bintt(prod, nx, ny, inverse=True), nx, ny


def poly_mul_naive(a, ax, ay, b, bx, by, nx, ny):
    out = [0] * (nx * ny)
    for i in range(ax):
        for j in range(ay):
            u = a[i * ay + j]
            if u == 0:
                continue
            for k in range(bx):
                for l in range(by):
                    v = b[k * by + l]
                    if v:
                        out[(i + k) * ny + (j + l)] = (out[(i + k) * ny + (j + l)] + u * v) % R_MOD
    return out


def div_by_vanishing_opt(p, x_size, y_size, c, d):
    """div_by_vanishing_opt (bivariate_polynomial/mod.rs:2284-2410) after optimize_size.

    p has shape x_size x y_size with x_size = m*c, y_size = n*d.  Returns
    (quo_x [x_size*y_size], quo_y [c*y_size]) with P = Q_X (X^c - 1) + Q_Y (Y^d - 1).
    """
    assert x_size % c == 0 and y_size % d == 0
    m = x_size // c
    acc = [0] * (c * y_size)
    for bx in range(m):
        for lx in range(c):
            for y in range(y_size):
                acc[lx * y_size + y] = (acc[lx * y_size + y] + p[(bx * c + lx) * y_size + y]) % R_MOD
    qy = [0] * (c * y_size)
    if y_size > d:
        for x in range(c):
            for y in range(y_size - d):
                prev = qy[x * y_size + y - d] if y >= d else 0
                qy[x * y_size + y] = (prev - acc[x * y_size + y]) % R_MOD
    b = list(p)
    if y_size > d:
        for x in range(c):
            for y in range(y_size - d):
                co = qy[x * y_size + y]
                b[x * y_size + y] = (b[x * y_size + y] + co) % R_MOD
                b[x * y_size + y + d] = (b[x * y_size + y + d] - co) % R_MOD
    qx = [0] * (x_size * y_size)
    if x_size > c:
        for x in range(x_size - c):
            for y in range(y_size):
                prev = qx[(x - c) * y_size + y] if x >= c else 0
                qx[x * y_size + y] = (prev - b[x * y_size + y]) % R_MOD
    return qx, qy


def div_by_ruffini(p, x_size, y_size, x, y):
    """div_by_ruffini (bivariate_polynomial/mod.rs:2412-2477).

    P = Q_X (X - x) + Q_Y (Y - y) + r.  Returns (q_x [x_size*y_size], q_y [y_size], r).
    """

    def ruff(v, pt):
        if len(v) < 2:
            return [0], v[0] % R_MOD
        n = len(v)
        q = [0] * n
        b = v[n - 1] % R_MOD
        q[n - 2] = b
        for i in range(3, n + 1):
            b = (v[n - i + 1] + b * pt) % R_MOD
            q[n - i] = b
        return q, (v[0] + b * pt) % R_MOD

    qx = [0] * (x_size * y_size)
    rx = [0] * y_size
    for j in range(y_size):
        q, rr = ruff(p[j::y_size], x)
        for i in range(x_size):
            qx[i * y_size + j] = q[i] if i < len(q) else 0
        rx[j] = rr
    qy, r = ruff(rx, y)
    qy = qy + [0] * (y_size - len(qy))
    return qx, qy, r


# ----------------------------------------------------------------------------
# G1: y^2 = x^3 + 4 over Fq.  Affine points are (x, y) tuples; None = identity.
# ----------------------------------------------------------------------------
def g1_is_on_curve(pt) -> bool:
    if pt is None:
        return True
    x, y = pt
    return (y * y - x * x * x - 4) % Q_MOD == 0


def g1_neg(pt):
    if pt is None:
        return None
    return (pt[0], (-pt[1]) % Q_MOD)


def g1_add(p1, p2):
    """Affine addition with all exceptional cases (G1serde Add, group_structures/mod.rs:895-910)."""
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, Q_MOD - 2, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, Q_MOD - 2, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    y3 = (lam * (x1 - x3) - y1) % Q_MOD
    return (x3, y3)


def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q_MOD
    B = Y * Y % Q_MOD
    C = B * B % Q_MOD
    D = 2 * ((X + B) * (X + B) - A - C) % Q_MOD
    E = 3 * A % Q_MOD
    F = E * E % Q_MOD
    X3 = (F - 2 * D) % Q_MOD
    Y3 = (E * (D - X3) - 8 * C) % Q_MOD
    Z3 = 2 * Y * Z % Q_MOD
    return (X3, Y3, Z3)


def _jac_add_affine(P, q):
    X1, Y1, Z1 = P
    if q is None:
        return P
    x2, y2 = q
    if Z1 == 0:
        return (x2, y2, 1)
    Z1Z1 = Z1 * Z1 % Q_MOD
    U2 = x2 * Z1Z1 % Q_MOD
    S2 = y2 * Z1 * Z1Z1 % Q_MOD
    H = (U2 - X1) % Q_MOD
    Rr = (S2 - Y1) % Q_MOD
    if H == 0:
        if Rr == 0:
            return _jac_double(P)
        return (1, 1, 0)
    HH = H * H % Q_MOD
    HHH = H * HH % Q_MOD
    V = X1 * HH % Q_MOD
    X3 = (Rr * Rr - HHH - 2 * V) % Q_MOD
    Y3 = (Rr * (V - X3) - Y1 * HHH) % Q_MOD
    Z3 = Z1 * H % Q_MOD
    return (X3, Y3, Z3)


def _jac_to_affine(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, Q_MOD - 2, Q_MOD)
    zi2 = zi * zi % Q_MOD
    return (X * zi2 % Q_MOD, Y * zi2 * zi % Q_MOD)


def g1_mul(pt, k: int):
    """G1serde Mul<ScalarField> (group_structures/mod.rs:929-947)."""
    k %= R_MOD
    if pt is None or k == 0:
        return None
    acc = (1, 1, 0)
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add_affine(acc, pt)
    return _jac_to_affine(acc)


def msm_naive(scalars, points):
    """msm::msm with MSMConfig::default() (call sites iotools/mod.rs:2093-2099,
    group_structures/mod.rs:108-114,135-141): sum_i s_i * P_i, result affine, identity = None."""
    assert len(scalars) == len(points)
    acc = None
    for s, pt in zip(scalars, points):
        acc = g1_add(acc, g1_mul(pt, s))
    return acc


def encode_poly(coeffs, x_size, y_size, xy_powers, rs_x, rs_y):
    """encode_poly_from_xy_powers_with_timing (iotools/mod.rs:2041-2113).

    xy_powers is the flat CRS grid, index rs_y*i + j <-> x^i y^j.
    """
    xd, yd = find_degree(coeffs, x_size, y_size)
    tx, ty = xd + 1, yd + 1
    if tx > rs_x or ty > rs_y:
        raise ValueError("Insufficient length of sigma.sigma_1.xy_powers")
    if tx * ty == 0:
        return None
    sc = resize_exact(coeffs, x_size, y_size, tx, ty)
    bases = [xy_powers[rs_y * i + j] for i in range(tx) for j in range(ty)]
    return msm_naive(sc, bases)


# ----------------------------------------------------------------------------
# Deterministic PRNG shared by oracle, tests and bench (SplitMix64 -> 32-byte
# draws reduced mod r).
# ----------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def fr(self) -> int:
        v = self.next() | (self.next() << 64) | (self.next() << 128) | (self.next() << 192)
        return v % R_MOD

    def frs(self, n: int):
        return [self.fr() for _ in range(n)]


def random_fr(seed: int, n: int):
    """Bit-identical to oracle.c:orc_random_fr -- element i has its own SplitMix64 stream."""
    out = []
    for i in range(n):
        g = SplitMix64(seed + i * 0x1000193)
        v = [g.next() for _ in range(4)]
        v[3] &= 0x7FFFFFFFFFFFFFFF
        x = v[0] | (v[1] << 64) | (v[2] << 128) | (v[3] << 192)
        while x >= R_MOD:
            x -= R_MOD
        out.append(x)
    return out
