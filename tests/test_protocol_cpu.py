"""Protocol layer on the host (no GPU): the restated setup -> prove0..4 -> preprocess -> verify flow runs end to end on the
oracle backend, the restated verifier accepts the proof and rejects tampered ones, the file formats round-trip, and the
pairing is pinned by the production verifier CRS embedded in the reference's browser verifier (tests/golden)."""
import copy
import json
import os

import pytest

import pyref as P
from tokamak_b200.protocol import formats as F
from tokamak_b200.protocol import fr
from tokamak_b200.protocol import pairing as PR
from tokamak_b200.protocol import preprocess as PP
from tokamak_b200.protocol import prover as PV
from tokamak_b200.protocol import setup as ST
from tokamak_b200.protocol import synthetic as S
from tokamak_b200.protocol import verifier as VF

HERE = os.path.dirname(os.path.abspath(__file__))


def _kat():
    d = json.load(open(os.path.join(HERE, "golden", "verifier_crs_kat.json")))["points"]
    g1 = lambda k: (int(d[k]["x"], 16), int(d[k]["y"], 16))
    g2 = lambda k: ((int(d[k]["x"][0], 16), int(d[k]["x"][1], 16)), (int(d[k]["y"][0], 16), int(d[k]["y"][1], 16)))
    return d, g1, g2


def test_production_crs_points_decode_on_curve():
    d, g1, g2 = _kat()
    for k, v in d.items():
        assert (P.g1_is_on_curve(g1(k)) if v["group"] == "G1" else PR.g2_is_on_curve(g2(k))), k
    assert g1("G") == P.G1_GEN and g2("H") == PR.G2_GEN


def test_pairing_kat_production_crs():
    """e(sigma1.x, H) = e(G, sigma2.x) and the same for y: real reference artefacts pin the pairing."""
    _, g1, g2 = _kat()
    assert PR.pairing_products_equal([g1("sigma1.x")], [g2("H")], [g1("G")], [g2("sigma2.x")])
    assert PR.pairing_products_equal([g1("sigma1.y")], [g2("H")], [g1("G")], [g2("sigma2.y")])
    assert not PR.pairing_products_equal([g1("sigma1.x")], [g2("H")], [g1("G")], [g2("sigma2.y")])


def test_pairing_bilinear_and_fixed_generators():
    assert P.g1_is_on_curve(ST.G1_FIXED) and PR.g2_is_on_curve(ST.G2_FIXED)
    assert PR.g2_mul(ST.G2_FIXED, PR.R) is None
    a, b = 0x1234567, 0x7654321
    assert PR.pairing_products_equal([P.g1_mul(ST.G1_FIXED, a)], [PR.g2_mul(ST.G2_FIXED, b)], [P.g1_mul(ST.G1_FIXED, a * b)], [ST.G2_FIXED])


def test_lagrange_bases_match_inverse_ntt_of_powers():
    """gen_evaled_lagrange_bases is defined through an inverse NTT of the power vector (vector_operations/mod.rs:19-28)."""
    val, size = 0xABCDEF0123456789, 16
    assert fr.lagrange_bases_at(val, size) == P.ntt(fr.powers(val, size), inverse=True)
    w = fr.root_of_unity(size)
    assert fr.lagrange_bases_at(pow(w, 3, fr.R_MOD), size) == [1 if k == 3 else 0 for k in range(size)]


def test_formats_roundtrip(tmp_path):
    params, infos, r1cs = S.make_library(S.tiny_shape())
    pl, perm, inst = S.synthesize(params, infos, r1cs)
    F.write_library(str(tmp_path / "lib"), params, infos, r1cs)
    F.write_synthesizer_output(str(tmp_path / "syn"), pl, perm, inst)
    assert F.read_library(str(tmp_path / "lib")) == (params, infos, r1cs)
    assert F.read_synthesizer_output(str(tmp_path / "syn")) == (pl, perm, inst)
    for p in pl:
        assert S.check_r1cs(r1cs[p.subcircuitId], p.variables)
    with open(tmp_path / "lib" / "r1cs" / "subcircuit0.r1cs", "r+b") as f:
        f.write(b"xxxx")
    with pytest.raises(ValueError):
        F.read_r1cs(str(tmp_path / "lib" / "r1cs" / "subcircuit0.r1cs"))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_r1cs_reader_on_the_reference_library():
    lib = "/root/reference/packages/frontend/qap-compiler/subcircuits/library"
    infos = json.load(open(os.path.join(lib, "subcircuitInfo.json")))
    for s in infos:
        r = F.read_r1cs(os.path.join(lib, "r1cs", f"subcircuit{s['id']}.r1cs"))
        assert (r.n_wires, r.n_constraints) == (s["Nwires"], s["Nconsts"])
    ref = json.load(open(os.path.join(lib, "setupParams.json")))
    mine, _, _ = S.make_library(S.reference_shape())
    for k in ("l_free", "l", "l_user_out", "l_user", "l_D", "n", "s_D", "s_max"):
        assert getattr(mine, k) == ref[k], k


def test_sparse_scalar_tables():
    """frs_sparse (the prover's vanishing / unit-vector / blinding tables without a Python loop over their zeros) against
    frs_from_ints on the dense list, values reduced mod r, later entries overriding nothing else."""
    import numpy as np

    from tokamak_b200 import frs_from_ints, frs_sparse

    dense = [0] * 64
    dense[0], dense[7], dense[63] = fr.R_MOD - 1, 5, (1 << 255) % fr.R_MOD
    assert np.array_equal(frs_sparse(64, {0: -1, 7: 5, 63: 1 << 255}), frs_from_ints(dense))
    assert not frs_sparse(8, {}).any()


def test_packed_real_library_fixture():
    """tests/golden/real_library.json.xz (the reference's circuit library packed by tests/golden/gen_real_library.py): shapes
    of setupParams.json / subcircuitInfo.json, and -- where the reference tree is present -- every constraint equal to the
    .r1cs binaries; a dataflow with a seeded witness keeps the reference's wire partition (O_pub_free = 109 public wires)."""
    here = os.path.dirname(os.path.abspath(__file__))
    params, infos, r1cs = F.read_packed_library(os.path.join(here, "golden", "real_library.json.xz"))
    params.validate()
    assert (params.l_free, params.l_user_out, params.l_user, params.l, params.l_D, params.m_D, params.n, params.s_D, params.s_max) == (128, 65, 85, 728, 4824, 26591, 4096, 14, 256)
    assert [s.name for s in infos[:5]] == ["bufferPubOut", "bufferPubIn", "bufferBlockIn", "bufferEVMIn", "bufferPrvIn"]
    assert sum(len(lc) for r in r1cs for abc in r.constraints for lc in abc) == 81624
    lib = "/root/reference/packages/frontend/qap-compiler/subcircuits/library"
    if os.path.isdir(lib):
        for s, r in zip(infos, r1cs):
            assert F.read_r1cs(os.path.join(lib, "r1cs", f"subcircuit{s.id}.r1cs")).constraints == r.constraints
    pl, perm, inst = S.synthesize(params, infos, r1cs, n_placements=12, seed=3, solver=S.fill_witness_unchecked(4))
    assert len(inst.a_pub_user) == params.l_user and len(inst.a_pub_function) == params.l - params.l_free
    assert len(inst.a_pub_user) + len(inst.a_pub_block) == 109  # O_pub_free has 109 points
    assert all(0 <= p.row < params.m_i and 0 <= p.X < params.m_i for p in perm)


@pytest.fixture(scope="module")
def tiny():
    from oracle_backend import OracleBackend

    be = OracleBackend()
    params, infos, r1cs = S.make_library(S.tiny_shape())
    pl, perm, inst = S.synthesize(params, infos, r1cs)
    sigma = ST.generate(be, params, infos, r1cs, ST.Tau.gen_fixed())
    pv = PV.Prover(be, params, infos, r1cs, sigma, pl, perm, inst, mixer=PV.Mixer.fixed(), checks=True)
    points, scalars, fmt, p4t = PV.prove(pv)
    pre = PP.preprocess(be, params, sigma, perm, inst)
    return dict(be=be, params=params, sigma=sigma, inst=inst, points=points, scalars=scalars, fmt=fmt, p4t=p4t, pre=pre, infos=infos, r1cs=r1cs, pl=pl, perm=perm)


def test_setup_identities(tiny):
    """The checks trusted-setup runs on its own output (setup/trusted-setup/src/main.rs:222-246): xy_powers[2 s_max] = x G,
    xy_powers[1] = y G, encode_poly(P) = P(x, y) G."""
    be, sigma, params = tiny["be"], tiny["sigma"], tiny["params"]
    tau = ST.Tau.gen_fixed()
    pts = sigma.xy_powers.points_host()
    from oracle_ffi import g1_to_tuple

    assert g1_to_tuple(pts[2 * params.s_max]) == P.g1_mul(ST.G1_FIXED, tau.x) == sigma.x
    assert g1_to_tuple(pts[1]) == P.g1_mul(ST.G1_FIXED, tau.y) == sigma.y
    import oracle_ffi as O

    coeffs = O.random_fr(5, 8 * 4)
    poly = be.from_coeffs(coeffs, 8, 4)
    assert be.commit(sigma.xy_powers, poly) == P.g1_mul(ST.G1_FIXED, poly.eval(tau.x, tau.y))


def test_full_snark_verifies_and_rejects_tampering(tiny):
    t = tiny
    assert VF.verify_arith(t["params"], t["sigma"], t["points"], t["scalars"], t["p4t"])
    assert VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], t["points"], t["scalars"])
    bad = dict(t["scalars"], V_eval=(t["scalars"]["V_eval"] + 1) % fr.R_MOD)
    assert not VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], t["points"], bad)
    badp = dict(t["points"], O_prv=t["be"].g1_add(t["points"]["O_prv"], t["sigma"].G))
    assert not VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], badp, t["scalars"])
    inst2 = copy.deepcopy(t["inst"])
    inst2.a_pub_user[0] = (inst2.a_pub_user[0] + 1) % fr.R_MOD
    assert not VF.verify_snark(t["params"], t["sigma"], t["pre"], inst2, t["points"], t["scalars"])


def test_verifier_rejects_malformed_inputs(tiny):
    """What the reference's deserialisation and indexing refuse (verify-rust/src/lib.rs:150-170 indexes a_pub_*[i] and panics;
    G1serde decoding never yields off-curve or non-canonical coordinates): a short instance, an off-curve proof point,
    a coordinate >= q, an evaluation >= r raise instead of flowing into the pairing."""
    t = tiny
    short = copy.deepcopy(t["inst"])
    short.a_pub_user = short.a_pub_user[:-1]
    with pytest.raises(ValueError):
        VF.verify_snark(t["params"], t["sigma"], t["pre"], short, t["points"], t["scalars"])
    x, y = t["points"]["U"]
    for bad_pt in ((x, (y + 1) % fr.Q_MOD), (x + fr.Q_MOD, y)):
        with pytest.raises(ValueError):
            VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], dict(t["points"], U=bad_pt), t["scalars"])
    with pytest.raises(ValueError):
        VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], t["points"], dict(t["scalars"], R_eval=t["scalars"]["R_eval"] + fr.R_MOD))


def test_unsatisfied_witness_is_caught(tiny):
    """A corrupted internal wire breaks the R1CS: the quotient identity check of prove0 (prove/src/lib.rs:1546-1556) fails."""
    t = tiny
    pl = copy.deepcopy(t["pl"])
    pl[5].variables[-1] = (pl[5].variables[-1] + 1) % fr.R_MOD
    pv = PV.Prover(t["be"], t["params"], t["infos"], t["r1cs"], t["sigma"], pl, t["perm"], t["inst"], mixer=PV.Mixer.fixed(), checks=True)
    with pytest.raises(AssertionError):
        pv.prove0()


def test_proof_json_layout(tiny):
    fmt = tiny["fmt"]
    assert len(fmt["proof_entries_part1"]) == 38 and len(fmt["proof_entries_part2"]) == 42
    assert all(len(s) == 2 + 32 for s in fmt["proof_entries_part1"]) and all(len(s) == 2 + 64 for s in fmt["proof_entries_part2"])
    assert F.recover_proof(fmt) == (tiny["points"], tiny["scalars"])
    pre = tiny["pre"]
    assert F.recover_preprocess(F.format_preprocess(pre)) == pre


def test_proof_is_deterministic_under_fixed_blinding(tiny):
    t = tiny
    pv = PV.Prover(t["be"], t["params"], t["infos"], t["r1cs"], t["sigma"], t["pl"], t["perm"], t["inst"], mixer=PV.Mixer.fixed())
    assert PV.prove(pv)[2] == t["fmt"]
    pv2 = PV.Prover(t["be"], t["params"], t["infos"], t["r1cs"], t["sigma"], t["pl"], t["perm"], t["inst"], mixer=PV.Mixer.fixed(seed=7))
    pts2, sc2, fmt2, _ = PV.prove(pv2)
    assert fmt2 != t["fmt"]
    assert VF.verify_snark(t["params"], t["sigma"], t["pre"], t["inst"], pts2, sc2)


@pytest.mark.parametrize("n_placements,small", [(6, 0.0), (7, 1.0)])
def test_fewer_placements_than_columns_and_random_blinding(tiny, n_placements, small):
    """Placement lists shorter than s_max leave empty columns (placement_variables.len() <= s_max, iotools/mod.rs:1308);
    random blinding scalars (Mixer.random) must verify like the fixed ones."""
    t = tiny
    pl, perm, inst = S.synthesize(t["params"], t["infos"], t["r1cs"], n_placements=n_placements, seed=9, small_value_fraction=small)
    pv = PV.Prover(t["be"], t["params"], t["infos"], t["r1cs"], t["sigma"], pl, perm, inst, mixer=PV.Mixer.random(), checks=True)
    points, scalars, _, _ = PV.prove(pv)
    pre = PP.preprocess(t["be"], t["params"], t["sigma"], perm, inst)
    assert VF.verify_snark(t["params"], t["sigma"], pre, inst, points, scalars)
    with pytest.raises(ValueError):
        PV.Prover(t["be"], t["params"], t["infos"], t["r1cs"], t["sigma"], pl * 2, perm, inst)  # more placements than s_max


def test_native_loaders_match_the_python_readers(tmp_path):
    """tkm_host_parse_r1cs / tkm_host_parse_hex_scalars (host-side data loaders of the library, no device needed) against the
    pure-Python readers: identical CSR arrays and witness values; malformed inputs are rejected with a status, not a crash."""
    import ctypes

    import numpy as np

    from tokamak_b200 import ffi
    from tokamak_b200.protocol import qap

    params, infos, r1cs = S.make_library(S.tiny_shape(), seed=11)
    pl, perm, inst = S.synthesize(params, infos, r1cs, seed=12, small_value_fraction=0.4)
    F.write_library(str(tmp_path / "lib"), params, infos, r1cs)
    F.write_synthesizer_output(str(tmp_path / "syn"), pl, perm, inst)
    p2, i2 = F.read_library_meta(str(tmp_path / "lib"))
    assert (p2, i2) == (params, infos)
    a, b = qap.library_csr_from_files(str(tmp_path / "lib"), p2, i2), qap.LibraryCSR(r1cs)
    for k in ("n_rows", "rp_base", "row_ptr", "wire", "coeff"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    npl, nperm, ninst = F.read_synthesizer_output(str(tmp_path / "syn"), infos)
    assert nperm == perm and ninst == inst and len(npl) == len(pl)
    for x, y in zip(npl, pl):
        assert x.subcircuitId == y.subcircuitId and list(x.variables) == y.variables and x.variables[2] == y.variables[2]
    wt_a, wt_b = qap.WitnessTable(params, npl, infos), qap.WitnessTable(params, pl, infos)
    assert np.array_equal(wt_a.values, wt_b.values) and np.array_equal(wt_a.var_off, wt_b.var_off)
    # one value moved from the first placement to the second keeps the total but shifts every later value: "Corrupted
    # placement variables" like the per-placement check of the reference (iotools/mod.rs:505-520), not a silent shift
    moved = copy.deepcopy(pl)
    moved[1].variables.insert(0, moved[0].variables.pop())
    F.write_synthesizer_output(str(tmp_path / "bad"), moved, perm, inst)
    with pytest.raises(ValueError, match="Corrupted placement variables"):
        F.read_synthesizer_output(str(tmp_path / "bad"), infos)
    lib = ffi.load()

    def parse(txt, cap=8):
        out = np.zeros((cap, 4), dtype=np.uint64)
        c = ctypes.c_size_t()
        rc = lib.tkm_host_parse_hex_scalars(txt, len(txt), out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(c))
        return rc, [int.from_bytes(out[i].tobytes(), "little") for i in range(c.value)]

    assert parse(b'{"k": ["0x0", "0x01", "0xFf", "0x%x", "0x%x"]}' % (fr.R_MOD - 1, fr.R_MOD + 5)) == (0, [0, 1, 255, fr.R_MOD - 1, 5])
    assert parse(b'["0xg1"]')[0] != 0 and parse(b'["0x1')[0] != 0 and parse(b'["0x' + b"f" * 65 + b'"]')[0] != 0 and parse(b'["0x1","0x2"]', cap=1)[0] != 0
    blob = open(tmp_path / "lib" / "r1cs" / "subcircuit5.r1cs", "rb").read()
    nw, nc, nnz = ctypes.c_uint32(), ctypes.c_uint32(), (ctypes.c_size_t * 3)()
    assert lib.tkm_host_parse_r1cs(blob, len(blob), ctypes.byref(nw), ctypes.byref(nc), nnz, None, None, None) == 0
    assert (nw.value, nc.value) == (infos[5].Nwires, infos[5].Nconsts)
    for bad in (b"xxxx" + blob[4:], blob[:-7], blob[:40]):
        assert lib.tkm_host_parse_r1cs(bad, len(bad), ctypes.byref(nw), ctypes.byref(nc), nnz, None, None, None) != 0
