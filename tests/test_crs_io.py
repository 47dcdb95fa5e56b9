"""CRS file formats on the host: the flat TZBWASM1 container of the browser prover
(packages/backend-wasm/src/artifacts/binary/binary-format.ts, binary-artifact-file.ts, specs/prover-crs.v1.json) and the rkyv
0.7 `SigmaRkyv` archive of the native prover (libs/src/iotools/mod.rs:1701-1783) -- header fields at the offsets the reference's
writer uses, the ffjavascript Montgomery point encoding pinned by the production verifier CRS embedded in the reference, and
write -> read round trips of a complete sigma."""
import hashlib
import json
import os
import re
import struct

import numpy as np
import pytest

import pyref as P
from oracle_backend import OracleBackend
from tokamak_b200.protocol import crs_io as C
from tokamak_b200.protocol import setup as ST
from tokamak_b200.protocol import synthetic as S

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tiny_sigma():
    be = OracleBackend()
    params, infos, r1cs = S.make_library(S.tiny_shape(), seed=3)
    return be, params, ST.generate(be, params, infos, r1cs, ST.Tau.gen_fixed())


def _same_sigma(a, b):
    for name in ("G", "H", "x", "y", "delta", "eta", "lagrange_KL", "delta_inv_alphak_xh_tx", "delta_inv_alpha4_xj_tx", "delta_inv_alphak_yi_ty"):
        assert getattr(a, name) == getattr(b, name), name
    assert a.sigma2 == b.sigma2
    for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
        ta, tb = getattr(a, name), getattr(b, name)
        assert (ta.rows, ta.cols) == (tb.rows, tb.cols), name
        assert np.array_equal(ta.points_host(), tb.points_host()), name


def test_ffjs_encoding_matches_the_production_verifier_crs():
    """The generator G embedded in packages/backend-wasm/src/verifier/generated/sigma-verify.generated.ts starts
    [22, 12, 83, 253, ...]: that is x * 2^384 mod q, little-endian -- the encoding this module writes."""
    b = C.g1_to_ffjs(P.G1_GEN)
    assert list(b[:4]) == [22, 12, 83, 253]
    assert C.g1_from_ffjs(b) == P.G1_GEN
    assert C.g1_to_ffjs(None) == bytes(96) and C.g1_from_ffjs(bytes(96)) is None
    kat = json.load(open(os.path.join(HERE, "golden", "verifier_crs_kat.json")))["points"]
    for name, v in kat.items():
        if v["group"] == "G1":
            pt = (int(v["x"], 16), int(v["y"], 16))
            assert C.g1_from_ffjs(C.g1_to_ffjs(pt)) == pt and C.g1_on_curve(pt), name
        else:
            pt = ((int(v["x"][0], 16), int(v["x"][1], 16)), (int(v["y"][0], 16), int(v["y"][1], 16)))
            assert C.g2_from_ffjs(C.g2_to_ffjs(pt)) == pt, name


def test_tzbwasm_header_layout_and_digest(tmp_path, tiny_sigma):
    be, params, sigma = tiny_sigma
    path = str(tmp_path / "prover_crs.bin")
    C.write_prover_crs(path, be, sigma, "tokamak-test/1.2.3")
    raw = open(path, "rb").read()
    # fixed header (binary-artifact-file.ts:57-75)
    assert raw[:8] == b"TZBWASM1"
    assert struct.unpack_from("<H", raw, 8)[0] == 1 and struct.unpack_from("<I", raw, 12)[0] == len(raw)
    kind_off, kind_len, ver_off, ver_len, dig_off, dig_len, sec_off, sec_len, data_off = struct.unpack_from("<9I", raw, 16)
    assert (kind_off, kind_len, ver_off, ver_len, dig_off, dig_len) == (64, 8, 72, 72, 144, 40)
    assert sec_off == 184 and sec_len == 9 * 96 and data_off == 184 + 9 * 96 and len(raw) % 8 == 0
    assert struct.unpack_from("<HH", raw, 52) == (9, 1)
    assert struct.unpack_from("<H", raw, kind_off)[0] == 6  # BinaryArtifactFileKind.ProverCrs
    assert raw[ver_off + 8:ver_off + 8 + struct.unpack_from("<H", raw, ver_off + 2)[0]] == b"tokamak-test/1.2.3"
    assert struct.unpack_from("<HH", raw, dig_off) == (1, 0xFFFF)
    zeroed = bytearray(raw)
    zeroed[dig_off + 8:dig_off + 40] = bytes(32)
    assert hashlib.sha256(zeroed).digest() == raw[dig_off + 8:dig_off + 40]
    # section table: the labels, types and encodings of specs/prover-crs.v1.json, in its order
    art = C.read_tzbwasm(path)
    assert art["kind"] == 6 and art["source_package_version"] == "tokamak-test/1.2.3"
    labels = [s["label"] for s in art["sections"]]
    assert labels == ["sigma.g1"] + C.G1_TABLE_LABELS + ["sigma.g2"]
    spec_path = "/root/reference/packages/backend-wasm/src/artifacts/specs/prover-crs.v1.json"
    if os.path.exists(spec_path):  # only in the build container: the spec itself
        spec = json.load(open(spec_path))
        assert [s["label"] for s in spec["sections"]] == labels
        assert [p["name"] for p in spec["sections"][0]["points"]] == C.G1_FIXED_NAMES
        assert [p["name"] for p in spec["sections"][-1]["points"]] == C.G2_FIXED_NAMES
    for s in art["sections"]:
        assert s["byte_offset"] % 8 == 0 and re.fullmatch(r"[a-z0-9][a-z0-9._-]*", s["label"])
        assert (s["type"], s["encoding"], s["element_bytes"]) == ((12, 4, 192) if s["label"] == "sigma.g2" else (11, 3, 96))
    assert art["sections"][1]["element_count"] == max(2 * params.n, 2 * params.m_i) * 2 * params.s_max
    # first point of xy_powers is x^0 y^0 G = G, in Montgomery form
    assert C.g1_from_ffjs(art["sections"][1]["data"][:96]) == sigma.G


def test_prover_crs_roundtrip_and_rejections(tmp_path, tiny_sigma):
    be, params, sigma = tiny_sigma
    path = str(tmp_path / "prover_crs.bin")
    C.write_prover_crs(path, be, sigma)
    _same_sigma(C.read_prover_crs(path, be, params), sigma)
    raw = bytearray(open(path, "rb").read())
    bad = str(tmp_path / "bad.bin")
    flipped = bytearray(raw)
    flipped[-100] ^= 1
    open(bad, "wb").write(flipped)
    with pytest.raises(ValueError, match="digest"):
        C.read_prover_crs(bad, be, params)
    open(bad, "wb").write(raw[:len(raw) - 8])
    with pytest.raises(ValueError):
        C.read_prover_crs(bad, be, params)
    open(bad, "wb").write(b"NOTMAGIC" + raw[8:])
    with pytest.raises(ValueError, match="TZBWASM1"):
        C.read_tzbwasm(bad)
    import copy
    other = copy.copy(params)
    other.s_max *= 2
    with pytest.raises(ValueError, match="setup parameters"):
        C.read_prover_crs(path, be, other)


@pytest.mark.parametrize("layout", C.LAYOUTS)
def test_sigma_rkyv_roundtrip_both_field_orders(tmp_path, tiny_sigma, layout):
    """The archived structs are repr(Rust): the reader must find the root object in either field order and follow the
    relative pointers of Vec and Vec<Vec<..>>."""
    be, params, sigma = tiny_sigma
    path = str(tmp_path / f"combined_sigma_{layout}.rkyv")
    C.write_sigma_rkyv(path, be, sigma, layout)
    got, detected = C.read_sigma_rkyv(path, be, params)
    assert detected == layout
    _same_sigma(got, sigma)
    raw = open(path, "rb").read()
    # root object at the very end; points are canonical little-endian [u8; 48] pairs (G1SerdeRkyv::from_g1serde)
    assert raw.count(C.g1_to_canonical(sigma.G)) >= 2  # xy_powers[0] and the G field
    trunc = str(tmp_path / "trunc.rkyv")
    open(trunc, "wb").write(raw[:-4])
    with pytest.raises(ValueError, match="Invalid sigma archive"):
        C.read_sigma_rkyv(trunc, be, params)
