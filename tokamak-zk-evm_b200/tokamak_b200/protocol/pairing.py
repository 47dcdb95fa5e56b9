"""BLS12-381 optimal-ate pairing on the host (pure Python integers).

Host-side CPU code like the reference's own: the verifier's pairing check is arkworks' `Bls12_381::multi_pairing`
on the CPU (libs/src/group_structures/mod.rs:120-124, called from verify-rust/src/lib.rs:243-289) and setup's G2
elements are a handful of scalar multiplications (`Sigma2::gen`, group_structures/mod.rs:752-777).  Nothing here is on
the data-parallel hot path.

Tower: Fq2 = Fq[u]/(u^2+1); Fq12 = Fq2[w]/(w^6 - xi), xi = 1 + u, elements are 6 Fq2 coefficients.  G2 lives on the
M-type twist y^2 = x^3 + 4 xi; untwist (x', y') -> (x' w^-2, y' w^-3).  `miller_product` omits the final conjugation for
the negative loop parameter: every pairing in a product is then the inverse of the standard one, which changes
neither equality checks between products nor bilinearity.
"""
Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000  # |x|, the curve parameter is -x
FINAL_EXP = (Q**12 - 1) // R

G2_GEN = (
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)


# ---------------------------------------------------------------- Fq2
def f2_add(a, b):
    return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def f2_neg(a):
    return ((-a[0]) % Q, (-a[1]) % Q)


def f2_mul(a, b):
    t0 = a[0] * b[0]
    t1 = a[1] * b[1]
    return ((t0 - t1) % Q, ((a[0] + a[1]) * (b[0] + b[1]) - t0 - t1) % Q)


def f2_sqr(a):
    return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)


def f2_scalar(a, k):
    return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], Q - 2, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


def f2_mul_xi(a):  # * (1 + u)
    return ((a[0] - a[1]) % Q, (a[0] + a[1]) % Q)


F2_ZERO = (0, 0)
F2_ONE = (1, 0)


# ---------------------------------------------------------------- Fq12 = Fq2[w]/(w^6 - xi)
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def f12_mul(a, b):
    t = [[0, 0] for _ in range(11)]
    for i in range(6):
        ai = a[i]
        if ai[0] == 0 and ai[1] == 0:
            continue
        a0, a1 = ai
        for j in range(6):
            b0, b1 = b[j]
            if b0 == 0 and b1 == 0:
                continue
            s = t[i + j]
            s[0] += a0 * b0 - a1 * b1
            s[1] += a0 * b1 + a1 * b0
    out = []
    for k in range(6):
        c0, c1 = t[k]
        if k < 5:
            h0, h1 = t[k + 6]
            c0 += h0 - h1  # * xi
            c1 += h0 + h1
        out.append((c0 % Q, c1 % Q))
    return tuple(out)


def f12_pow(a, e):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_mul(r, r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


# ---------------------------------------------------------------- G2 (affine on the twist, None = identity)
B2 = (4, 4)  # 4 * xi


def g2_is_on_curve(p):
    if p is None:
        return True
    x, y = p
    return f2_sqr(y) == f2_add(f2_mul(f2_sqr(x), x), B2)


def g2_neg(p):
    return None if p is None else (p[0], f2_neg(p[1]))


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if p[1] != q[1] or p[1] == F2_ZERO:
            return None
        lam = f2_mul(f2_scalar(f2_sqr(p[0]), 3), f2_inv(f2_scalar(p[1], 2)))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x3 = f2_sub(f2_sub(f2_sqr(lam), p[0]), q[0])
    return (x3, f2_sub(f2_mul(lam, f2_sub(p[0], x3)), p[1]))


def g2_mul(p, k):
    k %= R
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, p)
    return acc


# ---------------------------------------------------------------- Miller loop
def _line(lam, t, px, py):
    """(line through psi(T) with twist-slope lam, evaluated at P) * w^3 = (lam x' - y') - lam xP w^2 + yP w^3."""
    c0 = f2_sub(f2_mul(lam, t[0]), t[1])
    c2 = f2_scalar(lam, (-px) % Q)
    return (c0, F2_ZERO, c2, (py % Q, 0), F2_ZERO, F2_ZERO)


def miller_loop(p, q):
    """Unreduced ate Miller function f_{|x|,Q}(P); p = (x, y) ints in G1, q = ((x0,x1),(y0,y1)) in G2."""
    if p is None or q is None:
        return F12_ONE
    px, py = p
    f = F12_ONE
    t = q
    for bit in bin(X_ABS)[3:]:
        lam = f2_mul(f2_scalar(f2_sqr(t[0]), 3), f2_inv(f2_scalar(t[1], 2)))
        f = f12_mul(f12_mul(f, f), _line(lam, t, px, py))
        x3 = f2_sub(f2_sqr(lam), f2_scalar(t[0], 2))
        t = (x3, f2_sub(f2_mul(lam, f2_sub(t[0], x3)), t[1]))
        if bit == "1":
            lam = f2_mul(f2_sub(q[1], t[1]), f2_inv(f2_sub(q[0], t[0])))
            f = f12_mul(f, _line(lam, t, px, py))
            x3 = f2_sub(f2_sub(f2_sqr(lam), t[0]), q[0])
            t = (x3, f2_sub(f2_mul(lam, f2_sub(t[0], x3)), t[1]))
    return f


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def multi_pairing(g1s, g2s):
    """prod_i e(P_i, Q_i) (Bls12_381::multi_pairing as used by libs::group_structures::pairing)."""
    assert len(g1s) == len(g2s)
    f = F12_ONE
    for p, q in zip(g1s, g2s):
        f = f12_mul(f, miller_loop(p, q))
    return final_exponentiation(f)


def pairing_products_equal(lhs_g1, lhs_g2, rhs_g1, rhs_g2):
    """prod e(L_i, Q_i) == prod e(R_j, Q'_j), with one final exponentiation: prod e(L_i,Q_i) * prod e(-R_j,Q'_j) == 1."""
    neg = [None if p is None else (p[0], (-p[1]) % Q) for p in rhs_g1]
    return multi_pairing(list(lhs_g1) + neg, list(lhs_g2) + list(rhs_g2)) == F12_ONE
