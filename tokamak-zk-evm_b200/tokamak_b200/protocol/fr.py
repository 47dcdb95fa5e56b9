"""Scalar-field helpers on Python integers (host side; a few hundred scalar operations per proof)."""
R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
Q_MOD = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
_ROU_2_32 = pow(5, (R_MOD - 1) >> 32, R_MOD)  # primitive 2^32-th root of unity (5-based convention)


def inv(a):
    """ScalarField::inv; inv(0) = 0 like the reference backend."""
    return pow(a % R_MOD, R_MOD - 2, R_MOD)


def root_of_unity(n):
    """ntt::get_root_of_unity::<ScalarField>(n) for n a power of two."""
    assert n >= 1 and n & (n - 1) == 0 and n <= 1 << 32
    return pow(_ROU_2_32, (1 << 32) // n, R_MOD)


def from_hex(s):
    """ScalarField::from_hex: big-endian hex with optional 0x prefix; values are reduced mod r."""
    s = s[2:] if s.startswith(("0x", "0X")) else s
    return int(s, 16) % R_MOD if s else 0


def to_hex(a):
    return hex(a % R_MOD)


def powers(base, count):
    out = [1] * count
    for i in range(1, count):
        out[i] = out[i - 1] * base % R_MOD
    return out


def lagrange_bases_at(val, size):
    """gen_evaled_lagrange_bases (libs/src/vector_operations/mod.rs:19-28): [L_k(val)]_k over the size-th roots of
    unity, L_k(val) = (1/size) sum_i (val / w^k)^i = (val^size - 1) w^k / (size (val - w^k))."""
    w = root_of_unity(size)
    t = (pow(val, size, R_MOD) - 1) % R_MOD
    ninv = inv(size)
    wk = powers(w, size)
    if t == 0:  # val is itself a root of unity
        return [1 if wk[k] == val % R_MOD else 0 for k in range(size)]
    # batch inversion of (val - w^k)
    den = [(val - x) % R_MOD for x in wk]
    pref = [1] * (size + 1)
    for i, d in enumerate(den):
        pref[i + 1] = pref[i] * d % R_MOD
    run = inv(pref[size])
    out = [0] * size
    for i in range(size - 1, -1, -1):
        dinv = run * pref[i] % R_MOD
        run = run * den[i] % R_MOD
        out[i] = t * wk[i] % R_MOD * ninv % R_MOD * dinv % R_MOD
    return out
