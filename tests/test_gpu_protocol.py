"""Full prove on the B200 through the protocol driver: the GPU backend and the oracle backend
must produce byte-identical proof.json contents under fixed blinding, and the restated verifier must accept."""
import numpy as np
import pytest

from tokamak_b200.protocol import formats as F
from tokamak_b200.protocol import preprocess as PP
from tokamak_b200.protocol import prover as PV
from tokamak_b200.protocol import setup as ST
from tokamak_b200.protocol import synthetic as S
from tokamak_b200.protocol import verifier as VF

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backends():
    import tokamak_b200 as T
    from oracle_backend import OracleBackend
    from tokamak_b200.protocol.backend import GpuBackend

    ctx = T.Context(0)
    yield GpuBackend(ctx), OracleBackend()
    ctx.close()


def _small_shape():
    return S.LibrarySpec(n=64, s_max=16, m_i=256, l_user_out=4, l_user=8, l_free=16, l=32, n_prv_in=10, compute=[
        S.ComputeSpec("ALU1", 2, 5, 40), S.ComputeSpec("Poseidon", 2, 6, 64), S.ComputeSpec("Accumulator", 1, 16, 9),
        S.ComputeSpec("DecToBit", 30, 2, 31)])


@pytest.mark.parametrize("shape", ["tiny", "small"])
def test_gpu_proof_is_byte_identical_to_oracle_proof(backends, shape):
    gpu, orc = backends
    spec = S.tiny_shape() if shape == "tiny" else _small_shape()
    params, infos, r1cs = S.make_library(spec, seed=3)
    pl, perm, inst = S.synthesize(params, infos, r1cs, seed=4, small_value_fraction=0.3)
    tau = ST.Tau.gen_fixed()
    out = {}
    for be in (gpu, orc):
        sigma = ST.generate(be, params, infos, r1cs, tau)
        pv = PV.Prover(be, params, infos, r1cs, sigma, pl, perm, inst, mixer=PV.Mixer.fixed(), checks=True)
        points, scalars, fmt, p4t = PV.prove(pv)
        pre = PP.preprocess(be, params, sigma, perm, inst)
        out[be.name] = (sigma, points, scalars, fmt, pre)
    sg, so = out["b200"][0], out["oracle"][0]
    for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
        assert np.array_equal(getattr(sg, name).points_host(), getattr(so, name).points_host()), name
    assert out["b200"][3] == out["oracle"][3], "proof.json differs between the GPU path and the oracle path"
    assert F.format_preprocess(out["b200"][4]) == F.format_preprocess(out["oracle"][4])
    sigma, points, scalars, fmt, pre = out["b200"]
    assert VF.verify_snark(params, sigma, pre, inst, points, scalars)
    bad = dict(scalars, R_eval=(scalars["R_eval"] + 1) % VF.R_MOD)
    assert not VF.verify_snark(params, sigma, pre, inst, points, bad)


def test_reference_shape_prove_verifies_and_matches_golden_hash(backends):
    """Full size (n = 4096, s_max = 256, m_I = 4096, 256 placements): the proof under fixed blinding must verify and its
    SHA-256 must equal the committed value (tests/golden/prove_reference_shape.json, produced by this repository's GPU path
    and unchanged across every kernel / driver optimisation since; the oracle cross-check at full size takes minutes on a
    CPU and is done on the shape / 4 in bench.py and on the small shapes above)."""
    import hashlib
    import json
    import os

    gpu, _ = backends
    params, infos, r1cs = S.make_library(S.reference_shape())
    pl, perm, inst = S.synthesize(params, infos, r1cs, small_value_fraction=0.5)
    sigma = ST.generate(gpu, params, infos, r1cs, ST.Tau.gen_fixed())
    pv = PV.Prover(gpu, params, infos, r1cs, sigma, pl, perm, inst, mixer=PV.Mixer.fixed())
    points, scalars, fmt, _ = PV.prove(pv)
    pre = PP.preprocess(gpu, params, sigma, perm, inst)
    assert VF.verify_snark(params, sigma, pre, inst, points, scalars)
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prove_reference_shape.json")))
    assert hashlib.sha256(json.dumps(fmt, sort_keys=True).encode()).hexdigest() == golden["proof_sha256"]


def test_r1cs_uvw_polys_against_host_products(backends):
    """tkm_r1cs_uvw_polys (device sparse R1CS x witness + inverse biNTT) against the literal per-placement dot products of
    read_R1CS_gen_uvwXY on Python integers, with empty columns; invalid metadata is rejected before any launch."""
    import ctypes

    import oracle_ffi as O
    import tokamak_b200 as T
    from tokamak_b200.protocol import qap

    gpu, orc = backends
    params, infos, r1cs = S.make_library(_small_shape(), seed=5)
    pl, _, _ = S.synthesize(params, infos, r1cs, n_placements=9, seed=6, small_value_fraction=0.2)  # columns 9..15 stay empty
    gpu.init_ntt_domain(params.n * params.s_max)
    csr, wt = qap.LibraryCSR(r1cs), qap.WitnessTable(params, pl, infos)
    got = gpu.uvw_polys(params, csr, wt)
    from oracle_backend import uvw_evals

    exp = uvw_evals(params, pl, r1cs)
    for g, e in zip(got, exp):
        assert g.shape == (params.n, params.s_max)
        assert np.array_equal(g.to_rou_evals(), e)
        assert np.array_equal(g.copy_coeffs(), O.bintt(e, params.n, params.s_max, True))
    bad = np.array(wt.sub_of_col, copy=True)
    bad[0] = len(infos)  # subcircuit id out of range
    u, v, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    vals = np.ascontiguousarray(wt.values)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = gpu.ctx.lib.tkm_r1cs_uvw_polys(gpu.ctx.h, len(csr.n_rows), vp(csr.n_rows), vp(csr.rp_base), vp(csr.row_ptr), csr.row_ptr.shape[0], vp(csr.wire),
                                        vp(csr.coeff), csr.wire.shape[0], vp(bad), vp(wt.var_off), vp(vals), vals.shape[0], params.n, params.s_max,
                                        ctypes.byref(u), ctypes.byref(v), ctypes.byref(w))
    assert rc == T.ffi.TKM_ERR_INVALID_ARGUMENT if hasattr(T.ffi, "TKM_ERR_INVALID_ARGUMENT") else rc != 0


def test_r1cs_uvw_polys_on_the_reference_library(backends):
    """The same kernel on the REAL constraint structure: the reference's 14-subcircuit library (packed fixture
    tests/golden/real_library.json.xz, made by tests/golden/gen_real_library.py from the in-tree .r1cs binaries; 81 624
    non-zeros, 470 distinct coefficients, rows of up to 131 terms) with 40 placements and a seeded witness.  The products
    A w, B w, C w do not depend on satisfiability, so the literal per-placement dot products are the oracle."""
    import os

    import oracle_ffi as O
    from tokamak_b200.protocol import formats as F
    from tokamak_b200.protocol import qap

    gpu, _ = backends
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_library.json.xz")
    params, infos, r1cs = F.read_packed_library(path)
    assert (params.m_D, params.s_D, params.l, params.n, params.s_max) == (26591, 14, 728, 4096, 256)
    pl, _, _ = S.synthesize(params, infos, r1cs, n_placements=40, seed=21, solver=S.fill_witness_unchecked(22))
    gpu.init_ntt_domain(params.n * params.s_max)
    csr, wt = qap.LibraryCSR(r1cs), qap.WitnessTable(params, pl, infos)
    got = gpu.uvw_polys(params, csr, wt)
    from oracle_backend import uvw_evals

    exp = uvw_evals(params, pl, r1cs)
    for g, e in zip(got, exp):
        assert g.shape == (params.n, params.s_max)
        assert np.array_equal(g.to_rou_evals(), e)
    assert np.array_equal(got[0].copy_coeffs(), O.bintt(exp[0], params.n, params.s_max, True))


def test_random_polynomial_programs_gpu_vs_oracle(backends):
    """Differential fuzz: random sequences of DensePolynomialExt operations (mismatched shapes, zero polynomials, scalar
    forms, monomial shifts, coefficient scalings, divisions, evaluations) on the GPU engine and on its CPU twin; the padded
    coefficient matrices must agree after every step."""
    import random

    import oracle_ffi as O

    gpu, orc = backends
    gpu.init_ntt_domain(1 << 16)
    rng = random.Random(20261018)

    def same(a, b):
        tx, ty = max(a.shape[0], b.shape[0]), max(a.shape[1], b.shape[1])

        def padded(coeffs, shape):
            m = np.zeros((tx, ty, 4), dtype=np.uint64)
            m[:shape[0], :shape[1]] = np.asarray(coeffs, dtype=np.uint64).reshape(shape[0], shape[1], 4)
            return m

        return np.array_equal(padded(a.copy_coeffs(), a.shape), padded(b.copy_coeffs(), b.shape))

    for prog in range(12):
        pool = []
        for k in range(3):
            x, y = 1 << rng.randrange(0, 6), 1 << rng.randrange(0, 5)
            co = O.random_fr(9000 + 10 * prog + k, x * y)
            if rng.random() < 0.2:
                co[rng.randrange(x * y):] = 0  # low-degree / partly zero
            if rng.random() < 0.1:
                co[:] = 0
            pool.append((gpu.from_coeffs(co, x, y), orc.from_coeffs(co, x, y)))
        for step in range(10):
            op = rng.choice(["add", "sub", "mul", "scale", "adds", "mono", "scx", "scy", "eval", "neg", "ruffini"])
            (ga, oa), (gb, ob) = rng.choice(pool), rng.choice(pool)
            s = rng.randrange(VF.R_MOD) if rng.random() < 0.8 else rng.choice([0, 1, VF.R_MOD - 1])
            if op == "add":
                r = (ga + gb, oa + ob)
            elif op == "sub":
                r = (ga - gb, oa - ob)
            elif op == "mul":
                if ga.shape[0] * gb.shape[0] > 256 or ga.shape[1] * gb.shape[1] > 256:
                    continue
                r = (ga * gb, oa * ob)
            elif op == "scale":
                r = (ga * s, oa * s)
            elif op == "adds":
                r = (ga + s, oa + s)
            elif op == "neg":
                r = (-ga, -oa)
            elif op == "mono":
                ex, ey = rng.randrange(0, 3), rng.randrange(0, 3)
                if (ga.shape[0] + ex) > 128 or (ga.shape[1] + ey) > 64:
                    continue
                r = (ga.mul_monomial(ex, ey), oa.mul_monomial(ex, ey))
            elif op == "scx":
                r = (ga.scale_coeffs_x(s), oa.scale_coeffs_x(s))
            elif op == "scy":
                r = (ga.scale_coeffs_y(s), oa.scale_coeffs_y(s))
            elif op == "eval":
                px, py = rng.randrange(VF.R_MOD), rng.randrange(VF.R_MOD)
                assert ga.eval(px, py) == oa.eval(px, py), (prog, step, op)
                continue
            else:  # ruffini
                if ga.shape[0] < 2 or ga.shape[1] < 2:
                    continue
                px, py = rng.randrange(VF.R_MOD), rng.randrange(VF.R_MOD)
                gq, oq = ga.div_by_ruffini(px, py), oa.div_by_ruffini(px, py)
                assert gq[2] == oq[2] and same(gq[0], oq[0]) and same(gq[1], oq[1]), (prog, step, op)
                continue
            assert same(r[0], r[1]), (prog, step, op, r[0].shape, r[1].shape)
            if r[0].shape[0] * r[0].shape[1] <= 4096:
                pool[rng.randrange(len(pool))] = r


@pytest.mark.skipif(not __import__("os").environ.get("TKM_RUN_SLOW"), reason="minutes of CPU time: set TKM_RUN_SLOW=1")
def test_reference_shape_oracle_proof_matches_golden_hash(backends):
    """The CPU oracle proves the full-size circuit (CRS tables downloaded from the GPU setup, which the small-shape test
    checks point by point against the oracle's own setup) and must reproduce the golden proof hash: this is the
    cross-check that makes tests/golden/prove_reference_shape.json an oracle-verified value, not only a self-consistent one."""
    import copy
    import hashlib
    import json
    import os

    from oracle_backend import OracleTable

    gpu, orc = backends
    params, infos, r1cs = S.make_library(S.reference_shape())
    pl, perm, inst = S.synthesize(params, infos, r1cs, small_value_fraction=0.5)
    sg = copy.copy(ST.generate(gpu, params, infos, r1cs, ST.Tau.gen_fixed()))
    for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
        t = getattr(sg, name)
        setattr(sg, name, OracleTable(t.points_host(), t.rows, t.cols))
    pv = PV.Prover(orc, params, infos, r1cs, sg, pl, perm, inst, mixer=PV.Mixer.fixed())
    _, _, fmt, _ = PV.prove(pv)
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prove_reference_shape.json")))
    assert hashlib.sha256(json.dumps(fmt, sort_keys=True).encode()).hexdigest() == golden["proof_sha256"]
    print("oracle prove spans:", {k: round(v, 2) for k, v in pv.t.spans.items() if "." not in k})


def test_prove_from_crs_files(backends, tmp_path):
    """f3: a CRS written in the reference's two containers (flat TZBWASM1 prover_crs, rkyv SigmaRkyv archive) and loaded back
    onto the device -- TZBWASM1 sections go up as they are (device layout = ffjs Montgomery bytes, tkm_crs_upload_mont) --
    gives the same tables and the same proof bytes as the generated CRS."""
    from tokamak_b200.protocol import crs_io as C

    gpu, _ = backends
    params, infos, r1cs = S.make_library(_small_shape(), seed=3)
    pl, perm, inst = S.synthesize(params, infos, r1cs, seed=4, small_value_fraction=0.3)
    sigma = ST.generate(gpu, params, infos, r1cs, ST.Tau.gen_fixed())
    _, _, fmt0, _ = PV.prove(PV.Prover(gpu, params, infos, r1cs, sigma, pl, perm, inst, mixer=PV.Mixer.fixed()))
    flat, arch = str(tmp_path / "prover_crs.bin"), str(tmp_path / "combined_sigma.rkyv")
    C.write_prover_crs(flat, gpu, sigma)
    C.write_sigma_rkyv(arch, gpu, sigma)
    # the flat container holds the device bytes verbatim, which are the host conversion of the canonical points
    assert np.array_equal(sigma.xy_powers.mont_bytes_host()[:64], C.canonical_points_to_mont(sigma.xy_powers.points_host()[:64]))
    for loaded in (C.read_prover_crs(flat, gpu, params), C.read_sigma_rkyv(arch, gpu, params)[0]):
        for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
            assert np.array_equal(getattr(loaded, name).points_host(), getattr(sigma, name).points_host()), name
        _, _, fmt, _ = PV.prove(PV.Prover(gpu, params, infos, r1cs, loaded, pl, perm, inst, mixer=PV.Mixer.fixed()))
        assert fmt == fmt0
