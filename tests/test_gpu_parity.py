"""GPU parity tests: the CUDA path (through the C-ABI of libtokamak_b200) against the CPU oracle on the
same seeded inputs, against the committed golden vectors, and -- at full sizes -- through
size-independent identities.  Bit-exact everywhere (integer arithmetic).
Modeled on packages/backend/libs/src/tests.rs (test names cited per test)."""
import numpy as np
import pytest

import oracle_ffi as O
import pyref as P
from util import frs, fr1, g1_tuple, g1s, golden, ints, pt_from_golden, to_ints

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    import tokamak_b200 as T

    return T


@pytest.fixture(scope="module")
def ctx(T):
    O.build()
    c = T.Context(0)
    c.init_ntt_domain_for_size(1 << 23)
    yield c
    c.close()


# ------------------------------------------------------------------------------------------ fields
def test_fr_vec_ops_edge_and_random(ctx, T):
    edge = [0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (1 << 256) % P.R_MOD, (1 << 32) - 1, 1 << 32, P.R_MOD >> 1]
    a = np.concatenate([frs(edge * len(edge)), O.random_fr(11, 4096)])
    b = np.concatenate([frs([e for e in edge for _ in edge]), O.random_fr(12, 4096)])
    for op, name in ((T.OP_ADD, "add"), (T.OP_SUB, "sub"), (T.OP_MUL, "mul")):
        got = ctx.vec_op_host(op, a, b)
        assert np.array_equal(got, O.fr_vec_op(name, a, b)), name
    got = ctx.vec_op_host(T.OP_DIV, a[:512], b[:512])
    exp = O.fr_vec_op("mul", a[:512], O.fr_vec_inv(b[:512]))
    assert np.array_equal(got, exp)


def test_fr_mul_zero_columns_of_the_reduction(ctx, T):
    """The Fr reduction feeds the carry of column 0 (E[0] + m = 2^32 [E[0] != 0]) through the flag into the first fused
    multiply-add of every row (ff.cuh reduce_row): operands whose MONTGOMERY forms are k * 2^(32 j) make whole runs of rows
    see E[0] = 0 (no carry) next to rows with E[0] = 0xffffffff.. patterns.  All pairs against Python integers."""
    rinv = pow(1 << 256, -1, P.R_MOD)
    ks = [1, 2, 0xFFFFFFFF, 0x80000000, 0xFFFFFFFE, 0x73EDA753, (1 << 64) - 1, (1 << 96) + 1]
    vals = [k * (1 << (32 * j)) * rinv % P.R_MOD for j in range(8) for k in ks]  # Montgomery form = k * 2^(32 j) mod r
    vals += [0, 1, P.R_MOD - 1, rinv, (P.R_MOD - 1) * rinv % P.R_MOD]
    a = frs([x for x in vals for _ in vals])
    b = frs([y for _ in vals for y in vals])
    got = ctx.vec_op_host(T.OP_MUL, a, b)
    exp = frs([x * y % P.R_MOD for x in vals for y in vals])
    assert np.array_equal(got, exp)
    assert np.array_equal(got, O.fr_vec_op("mul", a, b))


def test_fr_vec_inv_batched(ctx):
    a = np.concatenate([frs([0, 1, P.R_MOD - 1, 0, 0, 7]), O.random_fr(13, 1001)])
    d = ctx.upload_fr(a)
    ctx.lib.tkm_fr_vec_inv(ctx.h, d, d, a.shape[0])
    got = ctx.download_fr(d, a.shape[0])
    ctx.dev_free(d)
    assert np.array_equal(got, O.fr_vec_inv(a))


def test_mont_roundtrip_and_root_of_unity(ctx):
    a = O.random_fr(14, 777)
    d = ctx.upload_fr(a)
    assert np.array_equal(ctx.download_fr(d, 777), a)
    ctx.dev_free(d)
    for k in (1, 2, 8, 12, 23, 32):
        assert ctx.get_root_of_unity(1 << k) == P.root_of_unity(1 << k)


# ------------------------------------------------------------------------------------------ NTT
def test_golden_bintt(ctx, T):
    g = golden()["bintt"]
    x, y, a = g["x"], g["y"], frs(ints(g["in"]))
    gx, gy = int(g["coset_x"], 16), int(g["coset_y"], 16)
    assert to_ints(ctx.bintt_host(a, x, y, T.FORWARD)) == ints(g["fwd"])
    assert to_ints(ctx.bintt_host(a, x, y, T.FORWARD, gx, gy)) == ints(g["fwd_coset"])
    assert to_ints(ctx.bintt_host(a, x, y, T.INVERSE)) == ints(g["inv"])
    assert to_ints(ctx.bintt_host(a, x, y, T.INVERSE, gx, gy)) == ints(g["inv_coset"])
    g1 = golden()["ntt_1d"]
    b = frs(ints(g1["in"]))
    assert to_ints(ctx.bintt_host(b, 16, 1, T.FORWARD)) == ints(g1["fwd_x16"])
    assert to_ints(ctx.bintt_host(b, 1, 16, T.INVERSE)) == ints(g1["inv_y16"])


SHAPES = [(1, 2), (2, 1), (2, 2), (4, 8), (64, 4), (16, 512), (1, 1024), (2048, 1), (4096, 1), (128, 1), (1, 256),
          (2048, 2), (4096, 16), (1024, 1024), (8192, 8), (16384, 4)]


@pytest.mark.parametrize("x,y", SHAPES)
def test_bintt_vs_oracle_all_modes(ctx, T, x, y):
    """_biNTT on every code path: single-pass / two-pass axes, degenerate axes, batch over rows or columns."""
    a = O.random_fr(100 + x + y, x * y)
    cx, cy = O.random_fr(5, 1)[0], O.random_fr(6, 1)[0]
    for inv in (False, True):
        for gx, gy in ((None, None), (cx, cy), (cx, None), (None, cy)):
            got = ctx.bintt_host(a, x, y, T.INVERSE if inv else T.FORWARD, gx, gy)
            exp = O.bintt(a, x, y, inv, gx, gy)
            assert np.array_equal(got, exp), (x, y, inv, gx is not None, gy is not None)


@pytest.mark.parametrize("x,y", [(4096, 256), (8192, 256), (8192, 512)])
def test_bintt_prover_shapes_vs_oracle(ctx, T, x, y):
    """The prover's transform shapes, full compare against the C oracle."""
    a = O.random_fr(200 + x, x * y)
    assert np.array_equal(ctx.bintt_host(a, x, y, T.FORWARD), O.bintt(a, x, y, False))
    assert np.array_equal(ctx.bintt_host(a, x, y, T.INVERSE), O.bintt(a, x, y, True))


def test_bintt_largest_shape_full_compare(ctx, T):
    """16384 x 512 = 2^23 (BASELINE.json config): every output element of the forward and inverse transform, plain and with
    cosets on both axes, against the C oracle (oracle.c does the full transform in a fraction of a second)."""
    x, y = 16384, 512
    a = O.random_fr(33, x * y)
    cx, cy = O.random_fr(35, 1)[0], O.random_fr(36, 1)[0]
    for inv in (False, True):
        for gx, gy in ((None, None), (cx, cy)):
            got = ctx.bintt_host(a, x, y, T.INVERSE if inv else T.FORWARD, gx, gy)
            assert np.array_equal(got, O.bintt(a, x, y, inv, gx, gy)), (inv, gx is not None)


def test_bintt_largest_shape_properties(ctx, T):
    """16384 x 512 = 2^23 (BASELINE.json config): round trip (tests.rs:107-131), coset == manual scaling
    (tests.rs:134-180), linearity, and sampled evaluations against Horner on the oracle."""
    x, y = 16384, 512
    n = x * y
    a = O.random_fr(31, n)
    d = ctx.upload_fr(a)
    e = ctx.dev_alloc(n * 32)
    ctx.bintt_dev(d, e, x, y, T.FORWARD)
    ev = ctx.download_fr(e, n)
    wx, wy = P.root_of_unity(x), P.root_of_unity(y)
    for (k, l) in ((0, 0), (1, 0), (0, 1), (12345, 77), (16383, 511), (8192, 256)):
        exp = O.eval_xy(a, x, y, fr1(pow(wx, k, P.R_MOD)), fr1(pow(wy, l, P.R_MOD)))
        assert np.array_equal(ev[k * y + l], exp), (k, l)
    ctx.bintt_dev(e, e, x, y, T.INVERSE)
    assert np.array_equal(ctx.download_fr(e, n), a)
    # coset forward == scale coefficients then plain forward
    gx, gy = 0x1234567 + (1 << 200), 0x7654321 + (1 << 199)
    ctx.bintt_dev(d, e, x, y, T.FORWARD, gx, gy)
    got = ctx.download_fr(e, n)
    scaled = O.scale_coeffs(a, x, y, fr1(gx), fr1(gy))
    ds = ctx.upload_fr(scaled)
    ctx.bintt_dev(ds, ds, x, y, T.FORWARD)
    assert np.array_equal(got, ctx.download_fr(ds, n))
    # inverse coset undoes it
    ctx.bintt_dev(e, e, x, y, T.INVERSE, gx, gy)
    assert np.array_equal(ctx.download_fr(e, n), a)
    for p in (d, e, ds):
        ctx.dev_free(p)


def test_ntt_batch_rows_vs_columns(ctx, T):
    """Row batch == column batch on the transpose (tests.rs:519-588,617-646)."""
    n, batch = 512, 64
    a = O.random_fr(41, n * batch)
    d = ctx.upload_fr(a)
    o = ctx.dev_alloc(n * batch * 32)
    ctx.ntt_batch_dev(d, o, n, batch, columns_batch=False)
    assert np.array_equal(ctx.download_fr(o, n * batch), O.ntt(a, n, batch, False))
    ctx.ntt_batch_dev(d, o, n, batch, columns_batch=True)
    assert np.array_equal(ctx.download_fr(o, n * batch), O.ntt(a, n, batch, True))
    ctx.dev_free(d)
    ctx.dev_free(o)


@pytest.mark.parametrize("x,y,world", [(64, 32, 4), (2048, 64, 2), (4096, 256, 8), (32, 16, 1), (16, 16, 16)])
def test_ntt_batch_scatter_single_gpu_emulation(ctx, T, x, y, world):
    """tkm_ntt_batch_scatter (fused multi-GPU re-sharding store) with all 'peers' on this GPU: after every emulated rank has
    run its row pass, peer p's buffer must hold the column shard [x][y/G] of the Y-transformed matrix; the inverse X pass
    scatters column shards back into row shards."""
    full = O.random_fr(300 + x, x * y)
    xl, yb = x // world, y // world
    exp_y = O.ntt(full, y, x, False, False).reshape(x, y, 4)  # transform along Y of every row
    peers = [ctx.dev_alloc(x * yb * 32) for _ in range(world)]
    for r in range(world):
        d = ctx.upload_fr(full.reshape(x, y, 4)[r * xl:(r + 1) * xl].reshape(-1, 4).copy())
        ctx.ntt_batch_scatter(d, y, xl, False, T.FORWARD, None, peers, 1, yb, r * xl)
        ctx.dev_free(d)
    for p in range(world):
        got = ctx.download_fr(peers[p], x * yb).reshape(x, yb, 4)
        assert np.array_equal(got, exp_y[:, p * yb:(p + 1) * yb]), (p, "row pass scatter")
    # inverse X pass over each column shard, scattered into row shards [x/G][y]; with a coset generator on the axis
    g = 0x1234567
    rows = [ctx.dev_alloc(xl * y * 32) for _ in range(world)]
    exp_x = O.ntt(np.ascontiguousarray(exp_y).reshape(-1, 4), x, y, True, True, O.fr_from_int(g)).reshape(x, y, 4)  # inverse along X of every column
    for r in range(world):
        ctx.ntt_batch_scatter(peers[r], x, yb, True, T.INVERSE, g, rows, y, 1, r * yb)
    for q in range(world):
        got = ctx.download_fr(rows[q], xl * y).reshape(xl, y, 4)
        assert np.array_equal(got, exp_x[q * xl:(q + 1) * xl]), (q, "column pass scatter")
    for p in peers + rows:
        ctx.dev_free(p)
    with pytest.raises(T.TkmError):
        ctx.ntt_batch_scatter(1, y, xl, False, T.FORWARD, None, [1, 2, 3], 1, yb, 0)  # peer count must be a power of two


def test_ntt_domain_errors(T):
    c = T.Context(0)
    a = O.random_fr(1, 16)
    with pytest.raises(T.TkmError) as e:  # bivariate_polynomial/mod.rs:1437-1439
        c.bintt_host(a, 4, 4)
    assert e.value.status == -3
    c.init_ntt_domain_for_size(8)
    with pytest.raises(T.TkmError) as e:  # bivariate_polynomial/mod.rs:1440-1445
        c.bintt_host(a, 4, 4)
    assert e.value.status == -3 and "too small" in str(e.value)
    c.init_ntt_domain_for_size(16)
    assert np.array_equal(c.bintt_host(a, 4, 4), O.bintt(a, 4, 4))
    with pytest.raises(T.TkmError):
        c.bintt_host(O.random_fr(1, 12), 3, 4)
    c.close()


# ------------------------------------------------------------------------------------------ G1 / MSM
def test_golden_msm(ctx):
    g = golden()["msm"]
    ss = frs(ints(g["scalars"]))
    pts = g1s([pt_from_golden(p) for p in g["points"]])
    assert g1_tuple(ctx.msm_g1_host(ss, pts)) == pt_from_golden(g["result"])


def test_g1_single_ops(ctx):
    a = g1s([P.G1_GEN])[0]
    k = 0x123456789ABCDEF123456789ABCDEF
    b = ctx.g1_mul(a, k)
    assert g1_tuple(b) == P.g1_mul(P.G1_GEN, k)
    assert g1_tuple(ctx.g1_add(a, b)) == P.g1_mul(P.G1_GEN, k + 1)
    assert g1_tuple(ctx.g1_add(a, a)) == P.g1_mul(P.G1_GEN, 2)
    assert g1_tuple(ctx.g1_add(a, g1s([P.g1_neg(P.G1_GEN)])[0])) is None
    assert g1_tuple(ctx.g1_add(a, g1s([None])[0])) == P.G1_GEN
    assert g1_tuple(ctx.g1_mul(a, 0)) is None
    assert g1_tuple(ctx.g1_mul(a, P.R_MOD - 1)) == P.g1_neg(P.G1_GEN)


def test_fixed_base_mul_vs_oracle(ctx):
    """N one-point MSMs == per-point scalar mul (cpu_and_gpu_versions_produce_same, tests.rs:19-45)."""
    G = g1s([P.G1_GEN_FIXED_TAU])[0]
    ks = np.concatenate([frs([0, 1, 2, P.R_MOD - 1]), O.random_fr(51, 252)])
    got = ctx.g1_fixed_base_mul(G, ks)
    assert np.array_equal(got, O.g1_fixed_base_mul_batch(G, ks))


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 127, 510, 1000, 4096])
def test_msm_random_affine_vs_oracle(ctx, n):
    """Random affine bases (not multiples of a known generator order): compare with the C oracle Pippenger."""
    G = g1s([P.G1_GEN])[0]
    pts = O.g1_fixed_base_mul_batch(G, O.random_fr(60 + n, n))
    ss = O.random_fr(61 + n, n)
    assert np.array_equal(ctx.msm_g1_host(ss, pts), O.msm_g1(ss, pts))


def test_msm_empty_and_degenerate(ctx):
    G = g1s([P.G1_GEN])[0]
    assert g1_tuple(ctx.msm_g1_host(np.zeros((0, 4), np.uint64), np.zeros((0, 12), np.uint64))) is None
    pts = O.g1_fixed_base_mul_batch(G, O.random_fr(71, 600))
    zeros = np.zeros((600, 4), dtype=np.uint64)
    assert g1_tuple(ctx.msm_g1_host(zeros, pts)) is None  # all-zero scalars
    ones = frs([1] * 600)
    exp = None
    for p in pts:
        exp = P.g1_add(exp, g1_tuple(p))
    assert g1_tuple(ctx.msm_g1_host(ones, pts)) == exp  # hot bucket: every digit identical
    same = np.tile(pts[0], (600, 1))  # all bases equal: every bucket add is a doubling candidate
    ss = O.random_fr(72, 600)
    assert np.array_equal(ctx.msm_g1_host(ss, same), O.msm_g1(ss, same))
    ident = np.zeros((600, 12), dtype=np.uint64)  # all bases identity
    assert g1_tuple(ctx.msm_g1_host(ss, ident)) is None
    # P and -P with the same scalar cancel
    pm = np.stack([pts[0], g1s([P.g1_neg(g1_tuple(pts[0]))])[0]])
    assert g1_tuple(ctx.msm_g1_host(frs([12345, 12345]), pm)) is None
    # scalar mix the witness polynomials have: zeros, ones, r-1, small values
    mix = frs(([0] * 7 + [1] * 5 + [P.R_MOD - 1] * 3 + [2, 3, 255, 65535, 65536]) * 30)
    assert np.array_equal(ctx.msm_g1_host(mix, pts), O.msm_g1(mix, pts))


LAMBDA_GLV = 0xAC45A4010001A40200000000FFFFFFFF  # phi(P) = lambda*P on G1; r = lambda^2 + lambda + 1 (csrc/glv.cuh)


def test_msm_glv_edge_scalars(ctx):
    """The GLV split inside k_decompose (k = k1 + k2*lambda, signed halves below 2^127) on the scalars that sit on its
    branch boundaries: halves that are zero, negative, maximal, multiples of lambda, and values at or above r (folded).
    One chain of the Horner tail being empty (all k2 = 0 / all k1 = 0) is a separate case."""
    G = g1s([P.G1_GEN])[0]
    lam, r = LAMBDA_GLV, P.R_MOD
    h1 = (lam + 1) // 2
    edge = [0, 1, 2, r - 1, r - 2, lam - 1, lam, lam + 1, lam // 2, lam // 2 + 1, h1 * lam, (h1 + 1) * lam, (h1 + 1) * lam - 1,
            (lam + 1) * lam, r // 2, r // 2 + 1, (1 << 127) - 1, 1 << 127, (1 << 128) - 1, 1 << 128, (1 << 255) % r, (1 << 64) - 1]
    edge += [q * lam + d for q in (1, h1 - 1, h1, h1 + 1, lam) for d in (0, 1, lam // 2, lam // 2 + 1, lam - 1)]
    edge = [e % r for e in edge]
    n = len(edge)
    ks = O.random_fr(81, n)
    pts = O.g1_fixed_base_mul_batch(G, ks)
    ki = to_ints(ks)
    assert g1_tuple(ctx.msm_g1_host(frs(edge), pts)) == P.g1_mul(P.G1_GEN, sum(a * b for a, b in zip(edge, ki)) % r)
    # every scalar separately against the scalar multiplication (one-point MSMs)
    for e, pt, k in list(zip(edge, pts, ki))[:24]:
        assert g1_tuple(ctx.msm_g1_host(frs([e]), pt[None, :])) == P.g1_mul(P.G1_GEN, e * k % r), hex(e)
    # only the k1 chain is populated (scalars below 2^100), then only the k2 chain (multiples of lambda)
    import random
    rng = random.Random(82)
    small = [rng.randrange(1 << 100) for _ in range(n)]
    assert g1_tuple(ctx.msm_g1_host(frs(small), pts)) == P.g1_mul(P.G1_GEN, sum(a * b for a, b in zip(small, ki)) % r)
    mult = [rng.randrange(1 << 100) * lam % r for _ in range(n)]
    assert g1_tuple(ctx.msm_g1_host(frs(mult), pts)) == P.g1_mul(P.G1_GEN, sum(a * b for a, b in zip(mult, ki)) % r)
    # non-canonical limbs (>= r) are folded into [0, r): same group element
    big = [r, r + 1, 2 * r + 5, (1 << 256) - 1]
    raw = np.array([[(v >> (64 * i)) & ((1 << 64) - 1) for i in range(4)] for v in big], dtype=np.uint64)
    exp = P.g1_mul(P.G1_GEN, sum((a % r) * b for a, b in zip(big, ki)) % r)
    assert g1_tuple(ctx.msm_g1_host(raw, pts[:4])) == exp


@pytest.mark.parametrize("n,pieces", [(1, 4), (5, 4), (1000, 3), (4097, 4), (6000, 16)])
def test_msm_host_pipelined_pieces(ctx, n, pieces, monkeypatch):
    """tkm_msm_g1_host cut into point ranges (copy/compute pipeline): every piece adds into the same bucket set, so the
    result must not depend on the cut.  Includes scalars that put every piece's digits into the same buckets."""
    G = g1s([P.G1_GEN])[0]
    pts = O.g1_fixed_base_mul_batch(G, O.random_fr(160 + n, n))
    ss = O.random_fr(161 + n, n)
    exp = O.msm_g1(ss, pts)
    monkeypatch.setenv("TKM_MSM_HOST_PIECES", str(pieces))
    assert np.array_equal(ctx.msm_g1_host(ss, pts), exp)
    hot = frs([0x0001000100010001000100010001000100010001000100010001000100010001 % P.R_MOD] * n)
    assert np.array_equal(ctx.msm_g1_host(hot, pts), O.msm_g1(hot, pts))
    monkeypatch.setenv("TKM_MSM_HOST_PIECES", "1")
    assert np.array_equal(ctx.msm_g1_host(ss, pts), exp)


@pytest.mark.parametrize("copy_bound", ["0", "1"])
def test_msm_host_pipelined_large(ctx, copy_bound, monkeypatch):
    """2^21 points through both automatic layouts of the host pipeline -- two pieces (a quarter, then the rest) when the call
    is compute-bound, four ending in a small one (2 : 3 : 2 : 1) when its copies dominate; scalars travel before bases and each
    pass waits for its bases inside.  Answer from known discrete logs (O(N) field work)."""
    monkeypatch.setenv("TKM_MSM_HOST_COPY_BOUND", copy_bound)
    n = 1 << 21
    G = g1s([P.G1_GEN])[0]
    ks = O.random_fr(170, n)
    ss = O.random_fr(171, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dpts = ctx.dev_alloc(n * 96)
    ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dpts)
    pts = np.empty((n, 12), dtype=np.uint64)
    ctx.d2h(pts, dpts)
    ctx.dev_free(dk)
    ctx.dev_free(dpts)
    assert np.array_equal(ctx.msm_g1_host(ss, pts), O.g1_mul(G, O.fr_inner_product(ss, ks)))


@pytest.mark.parametrize("logn", [16, 18, 20, 22, 24])
def test_msm_large_known_discrete_logs(ctx, logn):
    """Bases k_i*G generated on the device; answer (sum s_i k_i)*G from O(N) field work on the oracle
    (SURVEY.md §8d config 2).  2^22 is BASELINE.json's headline size, 2^24 the top of its sweep."""
    n = 1 << logn
    G = g1s([P.G1_GEN])[0]
    ks = O.random_fr(80 + logn, n)
    ss = O.random_fr(81 + logn, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dpts = ctx.dev_alloc(n * 96)
    ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dpts)
    # spot-check the generated bases against the oracle
    pts_head = np.empty((8, 12), dtype=np.uint64)
    ctx.d2h(pts_head, dpts)
    assert np.array_equal(pts_head, O.g1_fixed_base_mul_batch(G, ks[:8]))
    ctx.lib.tkm_g1_bases_to_mont(ctx.h, dpts, dpts, n)
    ds = ctx.upload_fr(ss, to_mont=False)
    got = ctx.msm_g1_dev(ds, False, dpts, n)
    exp = O.g1_mul(G, O.fr_inner_product(ss, ks))
    assert np.array_equal(got, exp)
    # Montgomery-form scalars give the same point
    ctx.lib.tkm_fr_to_mont(ctx.h, ds, ds, n)
    assert np.array_equal(ctx.msm_g1_dev(ds, True, dpts, n), exp)
    for p in (dk, dpts, ds):
        ctx.dev_free(p)


@pytest.mark.parametrize("kind", ["all_equal", "mostly_zero", "small_values", "extremes", "one_hot_window"])
def test_msm_skewed_scalar_distributions_large(ctx, kind):
    """2^17 points with the scalar shapes witness polynomials have (hot buckets spanning thousands of accumulation chunks,
    mostly-zero digits, values near r): answer from known discrete logs, (sum s_i k_i) G."""
    n = 1 << 17
    G = g1s([P.G1_GEN])[0]
    ks = O.random_fr(500, n)
    rng = np.random.default_rng(501)
    if kind == "all_equal":
        vals = [0x1D2C3B4A5968778695A4B3C2D1E0F0E1D2C3B4A5968778695A4B3C2D1E0F % P.R_MOD] * n
    elif kind == "mostly_zero":
        vals = [0] * n
        for i in rng.choice(n, size=n // 50, replace=False):
            vals[int(i)] = int(rng.integers(1, 1 << 62))
    elif kind == "small_values":
        vals = [int(v) for v in rng.integers(0, 1 << 16, size=n)]
    elif kind == "extremes":
        pool = [0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (1 << 255) % P.R_MOD, (1 << 128) - 1, 1 << 128]
        vals = [pool[int(i)] for i in rng.integers(0, len(pool), size=n)]
    else:  # every scalar has a single non-zero 16-bit window, the same one
        vals = [int(v) << 96 for v in rng.integers(1, 1 << 16, size=n)]
    ss = frs(vals)
    dk = ctx.upload_fr(ks, to_mont=False)
    dpts = ctx.dev_alloc(n * 96)
    ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dpts)
    ctx.lib.tkm_g1_bases_to_mont(ctx.h, dpts, dpts, n)
    ds = ctx.upload_fr(ss, to_mont=False)
    got = ctx.msm_g1_dev(ds, False, dpts, n)
    assert np.array_equal(got, O.g1_mul(G, O.fr_inner_product(ss, ks))), kind
    for p_ in (dk, dpts, ds):
        ctx.dev_free(p_)


def test_msm_rect_and_indexed(ctx):
    """Strided rectangle of a CRS grid (encode_poly, iotools/mod.rs:2061-2088) and sparse gather
    (msm_g1_bases over gathered rows, group_structures/mod.rs:266-300)."""
    G = g1s([P.G1_GEN])[0]
    rs_x, rs_y = 32, 16
    grid = O.g1_fixed_base_mul_batch(G, O.random_fr(91, rs_x * rs_y))
    sx, sy = 16, 16  # scalar matrix shape
    sc = O.random_fr(92, sx * sy)
    dg = ctx.upload_bases(grid)
    ds = ctx.upload_fr(sc)
    for rows, cols in ((16, 16), (13, 9), (1, 16), (16, 1), (5, 7)):
        got = ctx.msm_g1_rect_dev(ds, True, sy, dg, rs_y, rows, cols)
        assert np.array_equal(got, O.msm_g1_rect(sc, sy, grid, rs_y, rows, cols)), (rows, cols)
    idx = np.array([(7 * k * k + 3) % (rs_x * rs_y) for k in range(200)], dtype=np.uint32)
    sc2 = O.random_fr(93, 200)
    ds2 = ctx.upload_fr(sc2, to_mont=False)
    di = ctx.dev_alloc(idx.nbytes)
    ctx.h2d(di, idx)
    got = ctx.msm_g1_indexed_dev(ds2, False, dg, di, 200)
    assert np.array_equal(got, O.msm_g1(sc2, grid[idx]))
    for p in (dg, ds, ds2, di):
        ctx.dev_free(p)


def test_msm_prover_rectangles_of_full_crs_grid(ctx):
    """The prover's real commitment extents -- 4097 x 257, 4097 x 511, 8192 x 511 -- as strided rectangles of a
    device-resident 8192 x 512 grid (encode_poly trims the polynomial to its degree and addresses xy_powers with the
    CRS row stride, iotools/mod.rs:2061-2088).  Bases k_ij*G generated on the device; expected (sum s_ij k_ij)*G."""
    rs_x, rs_y = 8192, 512
    n = rs_x * rs_y
    G = g1s([P.G1_GEN])[0]
    ks = O.random_fr(610, n)
    ss = O.random_fr(611, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dg = ctx.dev_alloc(n * 96)
    ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dg)
    ctx.lib.tkm_g1_bases_to_mont(ctx.h, dg, dg, n)
    ds = ctx.upload_fr(ss, to_mont=False)
    k2, s2 = ks.reshape(rs_x, rs_y, 4), ss.reshape(rs_x, rs_y, 4)
    for rows, cols in ((4097, 257), (4097, 511), (8192, 511), (8192, 512)):
        got = ctx.msm_g1_rect_dev(ds, False, rs_y, dg, rs_y, rows, cols)
        sk = np.ascontiguousarray(k2[:rows, :cols]).reshape(-1, 4)
        sv = np.ascontiguousarray(s2[:rows, :cols]).reshape(-1, 4)
        assert np.array_equal(got, O.g1_mul(G, O.fr_inner_product(sv, sk))), (rows, cols)
    for p_ in (dk, dg, ds):
        ctx.dev_free(p_)


# ------------------------------------------------------------------------------------------ polynomial engine
def poly_from(T, ctx, a, x, y):
    return T.DensePolynomialExt.from_coeffs(ctx, a, x, y)


def test_golden_poly_ops(ctx, T):
    g = golden()
    x, y = g["bintt"]["x"], g["bintt"]["y"]
    a = frs(ints(g["bintt"]["in"]))
    p = poly_from(T, ctx, a, x, y)
    px, py = ints(g["poly"]["point"])
    assert p.eval(px, py) == int(g["poly"]["eval"], 16)
    assert p._scale(px, py).coeffs_ints() == ints(g["poly"]["scale"])
    qx, qy, r = p.div_by_ruffini(px, py)
    assert qx.coeffs_ints() == ints(g["poly"]["ruffini_qx"]) and qy.coeffs_ints() == ints(g["poly"]["ruffini_qy"])
    assert r == int(g["poly"]["ruffini_r"], 16)
    v = g["vanishing"]
    vqx, vqy = p.clone().div_by_vanishing_opt(v["c"], v["d"])
    assert vqx.coeffs_ints() == ints(v["qx"]) and vqy.coeffs_ints() == ints(v["qy"])
    m = p * p
    assert m.shape == (g["mul_self"]["nx"], g["mul_self"]["ny"]) and m.coeffs_ints() == ints(g["mul_self"]["out"])
    grid = g1s([pt_from_golden(q) for q in g["commit"]["grid"]])
    sigma = T.Sigma1(ctx, grid, 8, 4)
    assert g1_tuple(sigma.encode_poly(p)) == pt_from_golden(g["commit"]["result"])


def test_from_evals_and_to_evals(ctx, T):
    """test_from_evals / test_coset_ntt_matches_manual_scaling (tests.rs:107-180)."""
    x, y = 64, 32
    ev = O.random_fr(301, x * y)
    p = T.DensePolynomialExt.from_rou_evals(ctx, ev, x, y)
    assert np.array_equal(p.copy_coeffs(), O.bintt(ev, x, y, True))
    assert np.array_equal(p.to_rou_evals(), ev)
    gx, gy = 11111, 22222
    assert np.array_equal(p.to_rou_evals(gx, gy), O.bintt(p.copy_coeffs(), x, y, False, fr1(gx), fr1(gy)))
    q = T.DensePolynomialExt.from_rou_evals(ctx, ev, x, y, gx, gy)
    assert np.array_equal(q.copy_coeffs(), O.bintt(ev, x, y, True, fr1(gx), fr1(gy)))


def test_add_sub_neg_scalar_mismatched_shapes(ctx, T):
    """test_add / test_sub / mismatched sizes / scalar ops / neg (tests.rs:183-403,711-797)."""
    a = O.random_fr(311, 8 * 4)
    b = O.random_fr(312, 4 * 16)
    pa, pb = poly_from(T, ctx, a, 8, 4), poly_from(T, ctx, b, 4, 16)
    ai, bi = to_ints(a), to_ints(b)

    def ref(op):
        out = [0] * (8 * 16)
        for i in range(8):
            for j in range(16):
                u = ai[i * 4 + j] if j < 4 else 0
                v = bi[i * 16 + j] if i < 4 else 0
                out[i * 16 + j] = op(u, v) % P.R_MOD
        return out

    s = pa + pb
    assert s.shape == (8, 16) and s.coeffs_ints() == ref(lambda u, v: u + v)
    assert (pa - pb).coeffs_ints() == ref(lambda u, v: u - v)
    assert (pb - pa).coeffs_ints() == ref(lambda u, v: v - u)
    assert (-pa).coeffs_ints() == [(-u) % P.R_MOD for u in ai]
    k = 0xDEADBEEF12345
    assert (pa * k).coeffs_ints() == [u * k % P.R_MOD for u in ai]
    assert (pa + k).coeffs_ints() == [(ai[0] + k) % P.R_MOD] + ai[1:]
    assert (pa - k).coeffs_ints() == [(ai[0] - k) % P.R_MOD] + ai[1:]


def test_resize_optimize_find_degree_mul_monomial(ctx, T):
    """test_resize / test_optimize_size / update_degree_general_case / test_mul_monomial (tests.rs:886-932,1011-1039,1312-1360)."""
    x, y = 16, 8
    ai = [0] * (x * y)
    for (i, j, v) in ((0, 0, 5), (3, 2, 7), (5, 1, 9)):
        ai[i * y + j] = v
    p = poly_from(T, ctx, frs(ai), x, y)
    assert p.find_degree() == (5, 2)
    q = p.clone()
    q.optimize_size()
    assert q.shape == (8, 4) and q.coeffs_ints() == P.resize(ai, x, y, 6, 3)[0]
    q.resize(3, 17)
    exp, nx, ny = P.resize(P.resize(ai, x, y, 6, 3)[0], 8, 4, 3, 17)
    assert q.shape == (nx, ny) == (4, 32) and q.coeffs_ints() == exp
    z = T.DensePolynomialExt.zero(ctx, 4, 4)
    assert z.find_degree() == (-1, -1) and z.is_zero()
    z.optimize_size()
    assert z.shape == (4, 4)
    m = p.mul_monomial(3, 5)
    exp, nx, ny = P.mul_monomial(ai, x, y, 3, 5)
    assert m.shape == (nx, ny) and m.coeffs_ints() == exp
    with pytest.raises(T.TkmError):
        T.DensePolynomialExt.from_coeffs(ctx, frs([1] * 12), 3, 4)


def test_eval_and_partial_evals(ctx, T):
    """test_eval / test_eval_x / test_eval_y (tests.rs:838-883)."""
    x, y = 256, 64
    a = O.random_fr(321, x * y)
    p = poly_from(T, ctx, a, x, y)
    px, py = 0xABCDEF0123456789, (1 << 250) + 12345
    assert p.eval(px, py) == O.fr_to_int(O.eval_xy(a, x, y, fr1(px), fr1(py)))
    ex = p.eval_x(px)
    assert ex.shape == (1, y)
    ai = to_ints(a)
    assert ex.coeffs_ints() == P.eval_x(ai, x, y, px)
    ey = p.eval_y(py)
    assert ey.shape == (x, 1) and ey.coeffs_ints() == P.eval_y(ai, x, y, py)


def test_scale_coeffs(ctx, T):
    x, y = 128, 32
    a = O.random_fr(331, x * y)
    p = poly_from(T, ctx, a, x, y)
    s = 0x1234567890ABCDEF
    assert np.array_equal(p.scale_coeffs_x(s).copy_coeffs(), O.scale_coeffs(a, x, y, fr1(s), None))
    assert np.array_equal(p.scale_coeffs_y(s).copy_coeffs(), O.scale_coeffs(a, x, y, None, fr1(s)))


def test_mul_polynomial(ctx, T):
    """test_mul_polynomial (tests.rs:1042-1088) + scalar fast paths (bivariate_polynomial/mod.rs:1867-1877)."""
    a = O.random_fr(341, 32 * 16)
    b = O.random_fr(342, 8 * 64)
    pa, pb = poly_from(T, ctx, a, 32, 16), poly_from(T, ctx, b, 8, 64)
    m = pa * pb
    exp, nx, ny = P.poly_mul(to_ints(a), 32, 16, to_ints(b), 8, 64)
    assert m.shape == (nx, ny) and m.coeffs_ints() == exp
    c = poly_from(T, ctx, frs([7] + [0] * 15), 4, 4)
    assert (pa * c).coeffs_ints() == [u * 7 % P.R_MOD for u in to_ints(a)]
    assert (c * c).shape == (1, 1) and (c * c).coeffs_ints() == [49]
    z = T.DensePolynomialExt.zero(ctx, 2, 2)
    assert (pa * z).is_zero()


@pytest.mark.parametrize("sa,sb", [((32, 1), (16, 8)), ((16, 8), (32, 1)), ((1, 16), (8, 32)), ((8, 32), (1, 16)), ((64, 1), (1, 32)),
                                   ((1, 32), (64, 1)), ((16, 1), (32, 1)), ((1, 8), (1, 64)), ((2048, 1), (1024, 4)), ((4096, 1), (4096, 64))])
def test_mul_with_a_univariate_factor(ctx, T, sa, sb):
    """_mul when a factor is X- or Y-univariate (K0(X) * d(X, Y), L(X) * L(Y), t(X) * q in prove2/prove4): the one-axis
    convolution path (transform along that axis only) gives the shape and coefficients of the general product, also when the
    univariate factor is stored with a padded second axis of zeros and on axes long enough for the two-pass NTT."""
    a = O.random_fr(343 + sa[0] + sb[1], sa[0] * sa[1])
    b = O.random_fr(344 + sa[1] + sb[0], sb[0] * sb[1])
    pa, pb = poly_from(T, ctx, a, *sa), poly_from(T, ctx, b, *sb)
    ctx.init_ntt_domain_for_size(1 << 20)
    m = pa * pb
    exp, nx, ny = P.poly_mul(to_ints(a), sa[0], sa[1], to_ints(b), sb[0], sb[1]) if sa[0] * sb[0] <= 4096 else (None, None, None)
    if exp is not None:
        assert m.shape == (nx, ny) and m.coeffs_ints() == exp
    else:  # large: the product at random points
        for k in range(4):
            x, y = 0x1234567 + k, (1 << 200) + 77 * k
            assert m.eval(x, y) == pa.eval(x, y) * pb.eval(x, y) % P.R_MOD
    # the same factor stored with a zero-padded second axis (degree 0 along it, size 4)
    if sa[1] == 1 and exp is not None:
        wide = np.zeros((sa[0], 4, 4), dtype=np.uint64)
        wide[:, 0, :] = a
        pw = poly_from(T, ctx, wide.reshape(-1, 4), sa[0], 4)
        m2 = pw * pb
        m2.resize(nx, ny)
        assert m2.coeffs_ints() == exp


@pytest.mark.parametrize("x,y,c,d", [(16, 16, 4, 4), (64, 32, 16, 8), (8192, 512, 4096, 256), (256, 64, 64, 64 // 2)])
def test_div_by_vanishing_opt(ctx, T, x, y, c, d):
    """test_div_by_vanishing_opt_basic (tests.rs:1224-1237): build P = Qx t_x + Qy t_y, divide, compare with the oracle
    restatement of the reference recurrences."""
    a = O.random_fr(351 + x, x * y)  # top row/column non-zero with overwhelming probability: optimize_size keeps the shape
    p = poly_from(T, ctx, a, x, y)
    qx, qy = p.div_by_vanishing_opt(c, d)
    eqx, eqy = O.div_by_vanishing_opt(a, x, y, c, d)
    assert qx.shape == (x, y) and qy.shape == (c, y)
    assert np.array_equal(qx.copy_coeffs(), eqx) and np.array_equal(qy.copy_coeffs(), eqy)


def test_div_by_vanishing_reconstructs(ctx, T):
    x, y, c, d = 32, 16, 8, 4
    rng = P.SplitMix64(77)
    qx0 = [rng.fr() if i < x - c else 0 for i in range(x) for j in range(y)]
    qy0 = [rng.fr() if j < y - d else 0 for i in range(c) for j in range(y)]
    tX = T.DensePolynomialExt.from_coeffs(ctx, frs([P.R_MOD - 1] + [0] * (c - 1) + [1] + [0] * (c - 1)), 2 * c, 1)
    tY = T.DensePolynomialExt.from_coeffs(ctx, frs([P.R_MOD - 1] + [0] * (d - 1) + [1] + [0] * (d - 1)), 1, 2 * d)
    pq = poly_from(T, ctx, frs(qx0), x, y) * tX + poly_from(T, ctx, frs(qy0), c, y) * tY
    pq.optimize_size()
    assert pq.shape == (x, y)
    qx, qy = pq.div_by_vanishing_opt(c, d)
    assert qx.coeffs_ints() == qx0 and qy.coeffs_ints() == qy0


@pytest.mark.parametrize("x,y", [(2, 2), (1, 8), (8, 1), (64, 32), (4096, 256)])
def test_div_by_ruffini(ctx, T, x, y):
    """test_div_by_ruffini (tests.rs:935-952)."""
    a = O.random_fr(361 + x, x * y)
    p = poly_from(T, ctx, a, x, y)
    px, py = 987654321987654321, 123456789123456789
    qx, qy, r = p.div_by_ruffini(px, py)
    eqx, eqy, er = O.div_by_ruffini(a, x, y, fr1(px), fr1(py))
    assert np.array_equal(qx.copy_coeffs(), eqx) and np.array_equal(qy.copy_coeffs(), eqy) and r == O.fr_to_int(er)
    assert r == p.eval(px, py)


def _long_division(num, den):
    """Schoolbook polynomial long division on Python integers (coefficients low to high)."""
    num, dd = list(num), max(i for i, c in enumerate(den) if c)
    linv = pow(den[dd], P.R_MOD - 2, P.R_MOD)
    quo = [0] * len(num)
    for i in range(len(num) - dd - 1, -1, -1):
        f = num[i + dd] * linv % P.R_MOD
        quo[i] = f
        for k in range(dd + 1):
            num[i + k] = (num[i + k] - f * den[k]) % P.R_MOD
    return quo, num


@pytest.mark.parametrize("y_dir", [False, True])
def test_divide_x_and_divide_y(ctx, T, y_dir):
    """divide_x / divide_y (tests.rs:955-1008): P = Q D + R at a random point, every line equal to integer long division,
    the constant-denominator path and the reference's panics."""
    x, y = (32, 256) if y_dir else (256, 32)
    pc = O.random_fr(440 + y_dir, x * y)
    p = poly_from(T, ctx, pc, x, y)
    dlen = 16
    dc = O.random_fr(442 + y_dir, dlen)
    dc[dlen - 3:] = 0  # degree 12 inside a 16-slot buffer
    den = poly_from(T, ctx, dc, 1, dlen) if y_dir else poly_from(T, ctx, dc, dlen, 1)
    q, r = p.divide_y(den) if y_dir else p.divide_x(den)
    assert q.shape == (x, y) and r.shape == (x, y)
    a, b = 0x1234567890ABCDEF % P.R_MOD, 0xFEDCBA0987654321 % P.R_MOD
    assert p.eval(a, b) == (q.eval(a, b) * den.eval(a, b) + r.eval(a, b)) % P.R_MOD
    pm, qm, rm = (np.array(to_ints(z), dtype=object).reshape(x, y) for z in (pc, q.copy_coeffs(), r.copy_coeffs()))
    dints = to_ints(dc)
    for line in (0, 1, (x if y_dir else y) - 1):
        num = list(pm[line, :]) if y_dir else list(pm[:, line])
        eq, er = _long_division(num, dints)
        assert (list(qm[line, :]) if y_dir else list(qm[:, line])) == eq
        assert (list(rm[line, :]) if y_dir else list(rm[:, line])) == er
    rd = r.find_degree()
    assert (rd[1] if y_dir else rd[0]) < 12
    # constant denominator: quotient = p / c, remainder = 0 (1 x 1)
    cden = poly_from(T, ctx, frs([7]), 1, 1)
    q2, r2 = p.divide_y(cden) if y_dir else p.divide_x(cden)
    assert np.array_equal(q2.copy_coeffs(), (p * pow(7, P.R_MOD - 2, P.R_MOD)).copy_coeffs()) and r2.shape == (1, 1) and r2.is_zero()
    wrong = poly_from(T, ctx, O.random_fr(444, 4 * 4), 4, 4)  # bivariate denominator
    with pytest.raises(T.TkmError, match="univariate"):
        p.divide_y(wrong) if y_dir else p.divide_x(wrong)
    with pytest.raises(T.TkmError, match="Numer.degree < Denom.degree"):
        (den.divide_y(p.get_univariate_polynomial_y(0)) if y_dir else den.divide_x(p.get_univariate_polynomial_x(0)))
    with pytest.raises(T.TkmError, match="Divide by zero"):
        p.divide_y(T.DensePolynomialExt.zero(ctx)) if y_dir else p.divide_x(T.DensePolynomialExt.zero(ctx))
    # get_univariate_polynomial_x / _y (tests.rs:800-836)
    col = p.get_univariate_polynomial_x(3)
    row = p.get_univariate_polynomial_y(5)
    assert col.shape == (x, 1) and row.shape == (1, y)
    assert to_ints(col.copy_coeffs()) == list(pm[:, 3]) and to_ints(row.copy_coeffs()) == list(pm[5, :])


def test_encode_poly_fixed_tau(ctx, T):
    """encode_poly(P) == P(tau_x, tau_y) * G (setup/trusted-setup/src/main.rs:222-246) on a 64 x 32 grid built
    on the device with the fixed-tau generator; trimmed-rectangle, zero-polynomial and too-small-CRS paths."""
    rs_x, rs_y = 64, 32
    tx, ty = P.TAU_FIXED["x"], P.TAU_FIXED["y"]
    G = g1s([P.G1_GEN_FIXED_TAU])[0]
    mon = [pow(tx, h, P.R_MOD) * pow(ty, i, P.R_MOD) % P.R_MOD for h in range(rs_x) for i in range(rs_y)]
    grid = ctx.g1_fixed_base_mul(G, frs(mon))
    sigma = T.Sigma1(ctx, grid, rs_x, rs_y)
    a = O.random_fr(371, 32 * 16)
    p = poly_from(T, ctx, a, 32, 16)
    exp = O.g1_mul(G, O.eval_xy(a, 32, 16, fr1(tx), fr1(ty)))
    assert np.array_equal(sigma.encode_poly(p), exp)
    # sparse polynomial in a larger buffer: only the (deg+1) rectangle is committed
    ai = [0] * (64 * 32)
    ai[0], ai[5 * 32 + 3], ai[17 * 32 + 9] = 3, 1, P.R_MOD - 1
    ps = poly_from(T, ctx, frs(ai), 64, 32)
    exp = O.g1_mul(G, O.eval_xy(frs(ai), 64, 32, fr1(tx), fr1(ty)))
    assert np.array_equal(sigma.encode_poly(ps), exp)
    assert g1_tuple(sigma.encode_poly(T.DensePolynomialExt.zero(ctx, 8, 8))) is None
    big = poly_from(T, ctx, O.random_fr(372, 128 * 2), 128, 2)
    with pytest.raises(T.TkmError) as e:
        sigma.encode_poly(big)
    assert "Insufficient length" in str(e.value)


def test_commit_begin_end_tickets(ctx, T):
    """tkm_poly_commit_begin / tkm_commit_end: queued commitments resolved out of order equal the synchronous ones; the
    polynomial may be dropped right after begin; the zero polynomial and ticket exhaustion behave."""
    G = g1s([P.G1_GEN])[0]
    rs_x, rs_y = 64, 32
    sigma = T.Sigma1(ctx, O.g1_fixed_base_mul_batch(G, O.random_fr(401, rs_x * rs_y)), rs_x, rs_y)
    shapes = [(64, 32), (32, 32), (8, 4), (1, 1), (64, 1)]
    polys = [T.DensePolynomialExt.from_coeffs(ctx, O.random_fr(410 + i, x * y), x, y) for i, (x, y) in enumerate(shapes)]
    exp = [sigma.encode_poly(p) for p in polys]
    tickets = [sigma.encode_poly_begin(p) for p in polys]
    polys = None  # dropped (freed on the stream) while the commitments are in flight
    zt = sigma.encode_poly_begin(T.DensePolynomialExt.zero(ctx, 4, 4))
    for i in reversed(range(len(tickets))):
        assert np.array_equal(sigma.encode_poly_end(tickets[i]), exp[i]), i
    assert g1_tuple(sigma.encode_poly_end(zt)) is None
    with pytest.raises(T.TkmError):
        sigma.encode_poly_end(tickets[0])  # already consumed
    one = T.DensePolynomialExt.from_coeffs(ctx, frs([5]), 1, 1)
    held = [sigma.encode_poly_begin(one) for _ in range(32)]
    with pytest.raises(T.TkmError):
        sigma.encode_poly_begin(one)  # 33rd ticket
    for t in held:
        sigma.encode_poly_end(t)
    sigma.close()


def test_poly_expr_leaf_over_roots(ctx, T):
    """PolyExpr.poly_over_roots(p, mx, my) = p(X / w_mx, Y / w_my) as a rotated read of p's leaf transform
    (TKM_PEX_LEAF_SHIFT): equal to the leaf of the explicitly scaled polynomial (scale_coeffs_x / _y by the inverse roots,
    what prove2 builds for r(X/w, Y) and r(X/w, Y/w)), in a product with another leaf, on several domains; malformed operands
    are rejected."""
    import ctypes

    E = T.PolyExpr
    ctx.init_ntt_domain_for_size(1 << 14)
    px, py = 16, 8
    p = poly_from(T, ctx, O.random_fr(840, px * py), px, py)
    q = poly_from(T, ctx, O.random_fr(841, px * py), px, py)
    for mx, my, dx, dy in ((16, 0, 64, 16), (16, 8, 64, 16), (0, 8, 32, 32), (4, 2, 32, 16), (64, 16, 64, 16)):
        s = p.clone()
        if mx:
            s = s.scale_coeffs_x(pow(ctx.get_root_of_unity(mx), -1, P.R_MOD))
        if my:
            s = s.scale_coeffs_y(pow(ctx.get_root_of_unity(my), -1, P.R_MOD))
        exp = E.sub(E.mul(E.poly(s), E.poly(q)), E.poly(s)).evaluate_fused_with_domain(dx, dy, ctx)
        expr = E.sub(E.mul(E.poly_over_roots(p, mx, my), E.poly(q)), E.poly_over_roots(p, mx, my))
        got = expr.evaluate_fused_with_domain(dx, dy, ctx)
        assert np.array_equal(got.copy_coeffs(), exp.copy_coeffs()), (mx, my, dx, dy)
        c = expr.evaluate_coeffs(ctx)
        c.resize(dx, dy)
        assert np.array_equal(c.copy_coeffs(), exp.copy_coeffs())
    with pytest.raises(ValueError):
        E.poly_over_roots(p, 128, 0).evaluate_fused_with_domain(64, 16, ctx)  # root of order 128 on a 64-point axis
    with pytest.raises(ValueError):
        E.poly_over_roots(p, 12, 0)
    hs = (ctypes.c_void_p * 1)(p.h)
    cs = frs([0])
    h = ctypes.c_void_p()
    for word in (T.PEX_LEAF_SHIFT | (1 | 0 << 4) << 8, T.PEX_LEAF_SHIFT | (0 | 9 << 4) << 8, T.PEX_LEAF_SHIFT | (0 | 1 << 16) << 8):
        prog = np.array([word], dtype=np.uint32)
        assert ctx.lib.tkm_polyexpr_eval(ctx.h, hs, 1, prog.ctypes.data_as(ctypes.c_void_p), 1, cs.ctypes.data_as(ctypes.c_void_p), 1, 64, 16, ctypes.byref(h)) != 0


def test_msm_begin_end_and_device_gather(ctx, T):
    """tkm_msm_g1_begin / tkm_msm_g1_indexed_begin (resolved by tkm_commit_end, buffers released right after begin) against the
    synchronous entry points and the oracle, queued out of order with empty inputs in between; tkm_fr_gather against numpy
    indexing, out-of-range indices rejected."""
    import ctypes

    G = g1s([P.G1_GEN])[0]
    n_tab = 3000
    pts = O.g1_fixed_base_mul_batch(G, O.random_fr(801, n_tab))
    d_tab = ctx.upload_bases(pts)
    lib, vp = ctx.lib, ctypes.c_void_p
    jobs = []
    for k, n in enumerate((1, 17, 2048, 2999)):
        ss = O.random_fr(810 + k, n)
        idx = np.random.default_rng(820 + k).integers(0, n_tab, size=n).astype(np.uint32)
        d_s, d_i = ctx.upload_fr(ss, to_mont=False), ctx.dev_alloc(n * 4)
        ctx.h2d(d_i, idx)
        t1, t2 = ctypes.c_int32(), ctypes.c_int32()
        T.check(lib.tkm_msm_g1_indexed_begin(ctx.h, vp(d_s), 0, vp(d_tab), vp(d_i), n, ctypes.byref(t1)))
        T.check(lib.tkm_msm_g1_begin(ctx.h, vp(d_s), 0, vp(d_tab), n, ctypes.byref(t2)))
        ctx.dev_free(d_s)
        ctx.dev_free(d_i)
        jobs.append((t1.value, O.msm_g1(ss, pts[idx]), t2.value, O.msm_g1(ss, pts[:n])))
    te = ctypes.c_int32()
    T.check(lib.tkm_msm_g1_begin(ctx.h, None, 0, None, 0, ctypes.byref(te)))  # empty MSM: the identity
    out = np.zeros(12, dtype=np.uint64)
    for t1, e1, t2, e2 in reversed(jobs):
        T.check(lib.tkm_commit_end(ctx.h, t2, out.ctypes.data_as(vp)))
        assert np.array_equal(out, e2)
        T.check(lib.tkm_commit_end(ctx.h, t1, out.ctypes.data_as(vp)))
        assert np.array_equal(out, e1)
    T.check(lib.tkm_commit_end(ctx.h, te.value, out.ctypes.data_as(vp)))
    assert not out.any()
    # device gather
    tab = O.random_fr(830, 1000)
    sel = np.random.default_rng(831).integers(0, 1000, size=5000).astype(np.uint32)
    d_t, d_sel, d_o = ctx.upload_fr(tab, to_mont=False), ctx.dev_alloc(5000 * 4), ctx.dev_alloc(5000 * 32)
    ctx.h2d(d_sel, sel)
    T.check(lib.tkm_fr_gather(ctx.h, vp(d_t), 1000, vp(d_sel), 5000, vp(d_o)))
    got = np.empty((5000, 4), dtype=np.uint64)
    ctx.d2h(got, d_o)
    assert np.array_equal(got, tab[sel])
    sel[77] = 1000
    ctx.h2d(d_sel, sel)
    assert lib.tkm_fr_gather(ctx.h, vp(d_t), 1000, vp(d_sel), 5000, vp(d_o)) != 0
    for p_ in (d_t, d_sel, d_o, d_tab):
        ctx.dev_free(p_)


def test_poly_expr_fused_matches_coefficients(ctx, T):
    """test_poly_expr_fused_matches_coefficients (tests.rs:1240-1276), plus a larger instance checked against the oracle."""
    for sx, sy, seed in ((2, 2, 400), (16, 8, 410)):
        polys = [poly_from(T, ctx, O.random_fr(seed + k, sx * sy), sx, sy) for k in range(5)]
        a, b, c, d, e = polys
        E = T.PolyExpr
        expr = E.weighted_sum([
            (7, E.mul_x_minus_one(E.sub(E.mul(E.poly(a), E.poly(b)), E.mul(E.poly(c), E.poly(d))))),
            (11, E.mul(E.sub(E.poly(a), E.scalar(1)), E.poly(e))),
        ])
        coeff_result = expr.evaluate_coeffs()
        fused_result = expr.evaluate_fused()
        rng = P.SplitMix64(seed)
        for _ in range(4):
            x, y = rng.fr(), rng.fr()
            assert coeff_result.eval(x, y) == fused_result.eval(x, y)
        # value check against big-int arithmetic at one point
        x, y = rng.fr(), rng.fr()
        ev = [P.eval_xy(to_ints(p.copy_coeffs()), sx, sy, x, y) for p in polys]
        exp = (7 * (x - 1) * (ev[0] * ev[1] - ev[2] * ev[3]) + 11 * (ev[0] - 1) * ev[4]) % P.R_MOD
        assert fused_result.eval(x, y) == exp
        # a bigger domain than needed gives the same polynomial (evaluate_fused_with_domain)
        big = expr.evaluate_fused_with_domain(4 * sx, 4 * sy)
        assert big.eval(x, y) == exp
        with pytest.raises(ValueError):
            expr.evaluate_fused_with_domain(sx, sy)


def test_polyexpr_program_entry_point(ctx, T):
    """tkm_polyexpr_eval: the prover's p_comb shape (prove/src/lib.rs:2110-2146) on a 128 x 64 domain against the coefficient-
    domain operators, shared leaves transformed once, constant-only / scale-by-one programs, and the reference's panics."""
    E = T.PolyExpr
    sx, sy = 32, 16
    r, g, f, r1, r2, KL, K0 = [poly_from(T, ctx, O.random_fr(900 + k, sx * sy), sx, sy) for k in range(7)]
    kappa = 0x1234567890ABCDEF1234567890ABCDEF % P.R_MOD
    rg = E.mul(E.poly(r), E.poly(g))
    p1 = E.mul(E.sub(E.poly(r), E.scalar(1)), E.poly(KL))
    p2 = E.mul_x_minus_one(E.sub(rg, E.mul(E.poly(r1), E.poly(f))))
    p3 = E.mul(E.poly(K0), E.sub(rg, E.mul(E.poly(r2), E.poly(f))))
    expr = E.weighted_sum([(1, p1), (kappa, p2), (kappa * kappa % P.R_MOD, p3)])
    l0 = ctx.launch_count()
    fused = expr.evaluate_fused_with_domain(128, 64)  # degree (31 + 62, 15 + 30): K0 * (r g - ..) is a triple product
    launches = ctx.launch_count() - l0
    # 11 leaf occurrences (degree bound: one find_degree each) but 7 distinct leaves (r, g, f are shared): 7 x (pad + two
    # NTT passes) + ONE expression kernel + the two passes of the inverse biNTT -- not one pass per DAG node
    assert launches <= 11 + 7 * 3 + 1 + 2, launches
    ref = expr.evaluate_coeffs()
    ref.resize(128, 64)
    assert fused.shape == (128, 64)
    assert np.array_equal(fused.copy_coeffs(), ref.copy_coeffs())
    assert to_ints(E.scalar(5).evaluate_fused_with_domain(2, 2, ctx).copy_coeffs()) == [5, 0, 0, 0]
    assert np.array_equal(E.scale(1, E.poly(r)).evaluate_fused_with_domain(sx, sy).copy_coeffs(), r.copy_coeffs())
    assert to_ints(E.weighted_sum([]).evaluate_fused_with_domain(1, 1, ctx).copy_coeffs()) == [0]
    with pytest.raises(ValueError):
        expr.evaluate_fused_with_domain(64, 64)  # too small for the degree
    with pytest.raises(ValueError):
        expr.evaluate_fused_with_domain(192, 64)  # not a power of two
    # raw entry point: malformed programs are rejected, not executed
    import ctypes
    hs = (ctypes.c_void_p * 1)(r.h)
    out = ctypes.c_void_p()
    one = frs([1])
    for prog in ([T.PEX_ADD], [T.PEX_LEAF | 3 << 8], [T.PEX_LEAF, T.PEX_LEAF], [99], [T.PEX_LEAF, T.PEX_SCALE | 7 << 8]):
        pr = np.array(prog, dtype=np.uint32)
        st = ctx.lib.tkm_polyexpr_eval(ctx.h, hs, 1, pr.ctypes.data_as(ctypes.c_void_p), len(prog), one.ctypes.data_as(ctypes.c_void_p), 1, sx, sy, ctypes.byref(out))
        assert st != 0, prog


def test_poly_lincomb_matches_chained_operators(ctx, T):
    """tkm_poly_lincomb (poly_comb!, prove/src/lib.rs:30-38 and the shifted helpers :48-124): k-ary, mixed shapes, monomial
    shifts, unit coefficients -- against the chain of scalar products, mul_monomial and additions, and against big-int sums."""
    shapes = [(8, 4), (4, 16), (16, 2), (1, 1), (8, 4)]
    polys = [poly_from(T, ctx, O.random_fr(950 + k, x * y), x, y) for k, (x, y) in enumerate(shapes)]
    cs = [int(v) for v in to_ints(O.random_fr(960, 5))]
    cs[3] = 1
    terms = [(cs[0], polys[0], 0, 0), (cs[1], polys[1], 1, 0), (cs[2], polys[2], 0, 3), (cs[3], polys[3], 0, 0), (cs[4], polys[4], 2, 1)]
    fused = T.DensePolynomialExt.lincomb(terms)
    chained = None
    for c, p_, sx, sy in terms:
        t = (p_.mul_monomial(sx, sy) if (sx or sy) else p_) * c
        chained = t if chained is None else chained + t
    assert fused.shape == chained.shape
    assert np.array_equal(fused.copy_coeffs(), chained.copy_coeffs())
    ox, oy = fused.shape
    exp = [0] * (ox * oy)
    for (c, p_, sx, sy), (x, y) in zip(terms, shapes):
        co = p_.coeffs_ints()
        for i in range(x):
            for j in range(y):
                exp[(i + sx) * oy + j + sy] = (exp[(i + sx) * oy + j + sy] + c * co[i * y + j]) % P.R_MOD
    assert fused.coeffs_ints() == exp
    # two-term, unshifted, no coefficients == operator +
    plain = T.DensePolynomialExt.lincomb([(1, polys[0]), (1, polys[1])])
    assert np.array_equal(plain.copy_coeffs(), (polys[0] + polys[1]).copy_coeffs())
    with pytest.raises(T.TkmError):
        T.DensePolynomialExt.lincomb([(1, polys[0])] * 17)


def test_two_contexts_in_one_process(T):
    """Per-context launch state (shared-memory opt-in, occupancy caches) and cudaSetDevice at every entry: two contexts --
    on two devices when the box has them, else on the same one -- used alternately, each with its own NTT domain, run the
    64 KiB-tile NTT path (axis 1024) and an MSM and agree with the oracle."""
    import torch

    O.build()
    dev_b = 1 if torch.cuda.device_count() > 1 else 0
    ca, cb = T.Context(0), T.Context(dev_b)
    try:
        ca.init_ntt_domain_for_size(1 << 12)
        cb.init_ntt_domain_for_size(1 << 20)  # first pass of a 2^20 axis uses L = 1024 tiles (64 KiB of dynamic shared memory)
        a = O.random_fr(970, 1024 * 4)
        b = O.random_fr(971, 1 << 20)
        G = g1s([P.G1_GEN])[0]
        pts = O.g1_fixed_base_mul_batch(G, O.random_fr(972, 300))
        ss = O.random_fr(973, 300)
        exp_msm = O.msm_g1(ss, pts)
        for _ in range(2):
            assert np.array_equal(ca.bintt_host(a, 1024, 4, T.FORWARD), O.bintt(a, 1024, 4, False))
            assert np.array_equal(cb.bintt_host(b, 1 << 20, 1, T.FORWARD), O.bintt(b, 1 << 20, 1, False))
            assert np.array_equal(cb.bintt_host(a, 1024, 4, T.INVERSE), O.bintt(a, 1024, 4, True))
            assert np.array_equal(ca.msm_g1_host(ss, pts), exp_msm)
            assert np.array_equal(cb.msm_g1_host(ss, pts), exp_msm)
    finally:
        ca.close()
        cb.close()


def test_transpose_fill_x_minus_one(ctx, T):
    """VecOps::transpose (vector_operations/mod.rs:139), device_vec_from_scalar and x_minus_one_evals
    (bivariate_polynomial/mod.rs:452-457,504-518)."""
    import ctypes

    rows, cols = 100, 37
    a = O.random_fr(420, rows * cols)
    d = ctx.upload_fr(a)
    o = ctx.dev_alloc(rows * cols * 32)
    ctx.transpose_dev(d, o, rows, cols)
    got = ctx.download_fr(o, rows * cols).reshape(cols, rows, 4)
    assert np.array_equal(got, a.reshape(rows, cols, 4).transpose(1, 0, 2))
    s = frs([123456789])
    T.check(ctx.lib.tkm_fr_vec_fill(ctx.h, s.ctypes.data, ctypes.c_void_p(o), 50))
    assert to_ints(ctx.download_fr(o, 50)) == [123456789] * 50
    x, y = 64, 8
    b = O.random_fr(421, x * y)
    db = ctx.upload_fr(b)
    T.check(ctx.lib.tkm_fr_mul_x_minus_one(ctx.h, ctypes.c_void_p(db), ctypes.c_void_p(db), x, y))
    w = P.root_of_unity(x)
    bi = to_ints(b)
    exp = [bi[i * y + j] * (pow(w, i, P.R_MOD) - 1) % P.R_MOD for i in range(x) for j in range(y)]
    assert to_ints(ctx.download_fr(db, x * y)) == exp
    for p in (d, o, db):
        ctx.dev_free(p)


def test_div_by_vanishing_legacy_coset_formulation(ctx, T):
    """div_by_vanishing (the coset-NTT formulation, tests.rs:1216-1222) on a numerator in the ideal: same quotients as
    div_by_vanishing_opt and as the oracle; the denominator cache is reused on the second call; the reference's panics."""
    c, d = 16, 8
    for qx_rows in (c, 2 * c):  # numerator x-size 2c (m = 2) and 4c (m = 4)
        qx0 = T.DensePolynomialExt.from_coeffs(ctx, O.random_fr(430 + qx_rows, qx_rows * d), qx_rows, d)
        qy0 = T.DensePolynomialExt.from_coeffs(ctx, O.random_fr(431, c * d), c, d)
        tx = T.DensePolynomialExt.from_coeffs(ctx, frs([P.R_MOD - 1] + [0] * (c - 1) + [1] + [0] * (c - 1)), 2 * c, 1)
        ty = T.DensePolynomialExt.from_coeffs(ctx, frs([P.R_MOD - 1] + [0] * (d - 1) + [1] + [0] * (d - 1)), 1, 2 * d)
        p = qx0 * tx + qy0 * ty
        cache = T.DivByVanishingCache(seed=5)
        for _ in range(2):
            gx, gy = p.clone().div_by_vanishing(c, d, cache)
            ox, oy = p.clone().div_by_vanishing_opt(c, d)
            xs, ys = p.shape
            ex, ey = O.div_by_vanishing_opt(p.copy_coeffs(), xs, ys, c, d)
            assert np.array_equal(gx.copy_coeffs(), ox.copy_coeffs()) and np.array_equal(gx.copy_coeffs(), ex)
            assert np.array_equal(gy.copy_coeffs(), oy.copy_coeffs()) and np.array_equal(gy.copy_coeffs(), ey)
        assert len(cache.denom_x_eval_inv) == 1 and len(cache.denom_y_eval_inv) == 1
    small = T.DensePolynomialExt.from_coeffs(ctx, O.random_fr(432, 8 * 4), 8, 4)
    with pytest.raises(ValueError):
        small.div_by_vanishing(16, 8)  # "The numerator must have grater degrees than denominators."
    with pytest.raises(ValueError):
        small.div_by_vanishing(3, 4)   # "The denominators must have degress as powers of two."


@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 4096, 100003, 1 << 20])
def test_suffix_product_scan(ctx, n):
    """prove1's recursion scan (prove/src/lib.rs:1858-1867): r[n-1] = 1, r[i] = r[i+1] * s[i+1]."""
    import ctypes

    a = O.random_fr(440 + (n % 97), n)
    if n > 3:
        a[n // 2] = frs([1])[0]
    d = ctx.upload_fr(a)
    T_ = __import__("tokamak_b200")
    T_.check(ctx.lib.tkm_fr_suffix_product(ctx.h, ctypes.c_void_p(d), ctypes.c_void_p(d), n))
    got = ctx.download_fr(d, n)
    ctx.dev_free(d)
    if n <= 4096:
        ai = to_ints(a)
        exp = [1] * n
        for i in range(n - 2, -1, -1):
            exp[i] = exp[i + 1] * ai[i + 1] % P.R_MOD
        assert to_ints(got) == exp
    else:
        # identity check at full size: r[i] = r[i+1] * s[i+1] for every i, and the boundary value
        assert to_ints(got[-1:]) == [1]
        lhs = got[:-1]
        rhs = O.fr_vec_op("mul", got[1:], a[1:])
        assert np.array_equal(lhs, rhs)


def test_div_by_ruffini_long_axis(ctx, T):
    """Segmented Ruffini path (x >= 256) on the largest prover shape."""
    x, y = 8192, 64
    a = O.random_fr(450, x * y)
    p = poly_from(T, ctx, a, x, y)
    px, py = 0x123456789ABCDEF0123, 0xFEDCBA9876543210FED
    qx, qy, r = p.div_by_ruffini(px, py)
    eqx, eqy, er = O.div_by_ruffini(a, x, y, fr1(px), fr1(py))
    assert np.array_equal(qx.copy_coeffs(), eqx) and np.array_equal(qy.copy_coeffs(), eqy) and r == O.fr_to_int(er)


@pytest.mark.parametrize("x,y", [(8192, 512), (2, 2048), (1 << 17, 2), (1, 1024), (512, 1), (4096, 1024)])
def test_div_by_ruffini_and_eval_all_paths(ctx, T, x, y):
    """Every launch path of div_by_ruffini and eval against the C oracle: the prover's largest shape (segmented X chains, the
    block-scan carries and the block-scan Y chain), a Y axis beyond one block (serial Y chain), an X axis with more than 1024
    segments (serial carries), degenerate axes, and the widest Y a single block takes."""
    a = O.random_fr(470 + (x % 97) + y, x * y)
    p = poly_from(T, ctx, a, x, y)
    px, py = 0x0F1E2D3C4B5A69788796A5B4C3D2E1F0, (1 << 254) + 0x1234567
    qx, qy, r = p.div_by_ruffini(px, py)
    eqx, eqy, er = O.div_by_ruffini(a, x, y, fr1(px), fr1(py))
    assert np.array_equal(qx.copy_coeffs(), eqx) and np.array_equal(qy.copy_coeffs(), eqy) and r == O.fr_to_int(er)
    assert p.eval(px, py) == O.fr_to_int(O.eval_xy(a, x, y, fr1(px), fr1(py))) == r
    assert p.eval(0, py) == O.fr_to_int(O.eval_xy(a, x, y, fr1(0), fr1(py)))
    assert p.eval(px, 0) == O.fr_to_int(O.eval_xy(a, x, y, fr1(px), fr1(0)))


def test_vector_operations_module(ctx, T):
    """libs/src/vector_operations helpers by their reference names (tests.rs:1487-1524,1595-1622)."""
    from tokamak_b200 import vector_operations as V

    n = 1000
    a, b = O.random_fr(460, n), O.random_fr(461, n)
    ai, bi = to_ints(a), to_ints(b)
    assert np.array_equal(V.point_mul_two_vecs(ctx, a, b), O.fr_vec_op("mul", a, b))
    assert np.array_equal(V.point_add_two_vecs(ctx, a, b), O.fr_vec_op("add", a, b))
    assert np.array_equal(V.point_div_two_vecs(ctx, a, b), O.fr_vec_op("mul", a, O.fr_vec_inv(b)))
    s = 0xABCDEF123456789
    assert to_ints(V.scale_vec(ctx, s, a)) == [s * u % P.R_MOD for u in ai]
    assert to_ints(V.scalar_vec_add(ctx, s, a)) == [(s + u) % P.R_MOD for u in ai]
    assert to_ints(V.scalar_vec_sub(ctx, s, a)) == [(s - u) % P.R_MOD for u in ai]
    assert V.inner_product_two_vecs(ctx, a, b) == O.fr_to_int(O.fr_inner_product(a, b))
    assert V.vec_sum(ctx, a) == sum(ai) % P.R_MOD
    prod = 1
    for u in ai[:300]:
        prod = prod * u % P.R_MOD
    assert V.vec_product(ctx, a[:300]) == prod
    op = V.outer_product_two_vecs(ctx, a[:13], b[:7])
    assert to_ints(op) == [ai[i] * bi[j] % P.R_MOD for i in range(13) for j in range(7)]
    tr = V.transpose_inplace(ctx, a[:15 * 20], 15, 20)
    assert np.array_equal(tr.reshape(20, 15, 4), a[:300].reshape(15, 20, 4).transpose(1, 0, 2))
    val = 0x1234567
    lag = V.gen_evaled_lagrange_bases(ctx, val, 64)
    assert to_ints(lag) == P.ntt([pow(val, i, P.R_MOD) for i in range(64)], inverse=True)
    assert to_ints(V.extend_monomial_vec(ctx, frs([1, 5, 25]), 6)) == [1, 5, 25, 125, 625, 3125]
    assert to_ints(V.resize(frs(list(range(6))), 2, 3, 3, 2)) == [0, 1, 3, 4, 0, 0]


def test_g1_sum(ctx):
    pts = [P.g1_mul(P.G1_GEN, k) for k in (3, 5, 7, 11)]
    arr = g1s(pts + [None, pts[0], P.g1_neg(pts[1])])
    exp = None
    for p in pts + [None, pts[0], P.g1_neg(pts[1])]:
        exp = P.g1_add(exp, p)
    assert g1_tuple(ctx.g1_sum(arr)) == exp
    assert g1_tuple(ctx.g1_sum(g1s([pts[2], P.g1_neg(pts[2])]))) is None
    assert g1_tuple(ctx.g1_sum(np.zeros((0, 12), dtype=np.uint64))) is None


@pytest.mark.parametrize("c", [8, 13, 16, 20])
def test_commit_with_fixed_base_tables(ctx, T, c):
    """tkm_crs_precompute: commitments through the fixed-base tables equal the plain ones and the oracle MSM."""
    G = g1s([P.G1_GEN])[0]
    rs_x, rs_y = 64, 32
    grid = O.g1_fixed_base_mul_batch(G, O.random_fr(470, rs_x * rs_y))
    grid[5] = 0  # an identity point inside the CRS
    sigma = T.Sigma1(ctx, grid, rs_x, rs_y)
    polys = []
    for (x, y, seed) in ((64, 32, 471), (32, 16, 472), (16, 32, 473)):
        a = O.random_fr(seed, x * y)
        a[::7] = 0
        a[3] = frs([P.R_MOD - 1])[0]
        a[4] = frs([1])[0]
        polys.append((a, x, y))
    plain = [sigma.encode_poly(poly_from(T, ctx, a, x, y)) for a, x, y in polys]
    sigma.precompute(c)
    for (a, x, y), ref in zip(polys, plain):
        got = sigma.encode_poly(poly_from(T, ctx, a, x, y))
        assert np.array_equal(got, ref)
        assert np.array_equal(got, O.msm_g1_rect(a, y, grid, rs_y, x, y))
    sparse = [0] * (64 * 32)
    sparse[0], sparse[40 * 32 + 9] = 5, P.R_MOD - 2
    exp = O.msm_g1_rect(frs(sparse), 32, grid, rs_y, 64, 32)
    assert np.array_equal(sigma.encode_poly(poly_from(T, ctx, frs(sparse), 64, 32)), exp)
    assert g1_tuple(sigma.encode_poly(T.DensePolynomialExt.zero(ctx, 8, 8))) is None


def test_reflected_scalar_operators(ctx, T):
    """&s + &p and &s - &p (bivariate_polynomial/mod.rs:1100-1281): the scalar only touches c00."""
    a = O.random_fr(450, 8 * 4)
    p = poly_from(T, ctx, a, 8, 4)
    s = 0x123456789
    ai = to_ints(a)
    assert to_ints((s + p).copy_coeffs()) == [(ai[0] + s) % P.R_MOD] + ai[1:]
    assert to_ints((s - p).copy_coeffs()) == [(s - ai[0]) % P.R_MOD] + [(-v) % P.R_MOD for v in ai[1:]]
