"""File formats either side of the path: the qap-compiler library (setupParams.json,
subcircuitInfo.json, globalWireList.json, r1cs/subcircuit{i}.r1cs), the synthesizer outputs (placementVariables.json,
permutation.json, instance.json) and the Solidity-split proof / preprocess JSON.  Field names are the reference's
(libs/src/iotools/mod.rs:167-177,367-372,400-416,459-467)."""
import json
import os
import struct
from dataclasses import asdict, dataclass, field
from typing import List

from .fr import Q_MOD, R_MOD, from_hex, to_hex


@dataclass
class SetupParams:
    l_free: int
    l: int
    l_user_out: int
    l_user: int
    l_D: int
    m_D: int
    n: int
    s_D: int
    s_max: int

    @property
    def m_i(self):
        return self.l_D - self.l

    def validate(self):
        """validate_setup_shape / validate_public_wire_size (libs/src/utils/mod.rs): NTT sizes must be powers of two."""
        for name, v in (("n", self.n), ("s_max", self.s_max), ("m_i", self.m_i), ("l_free", self.l_free)):
            if v < 1 or v & (v - 1):
                raise ValueError(f"{name} = {v} must be a power of two")
        if not (self.l_user_out <= self.l_user <= self.l_free <= self.l <= self.l_D <= self.m_D):
            raise ValueError("inconsistent wire partition in setup parameters")


@dataclass
class SubcircuitInfo:
    id: int
    name: str
    Nwires: int
    Nconsts: int
    Out_idx: List[int]
    In_idx: List[int]
    flattenMap: List[int]


class ScalarArray:
    """A placement's variables as a (k, 4) uint64 array of canonical little-endian limbs (what the native loader
    produces and what the device takes) that still reads like a list of integers for the host-side code paths."""

    def __init__(self, limbs):
        import numpy as np

        self.limbs = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)

    def __len__(self):
        return self.limbs.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        a = self.limbs[i]
        return int(a[0]) | (int(a[1]) << 64) | (int(a[2]) << 128) | (int(a[3]) << 192)

    def __iter__(self):
        b = self.limbs.tobytes()
        return (int.from_bytes(b[k:k + 32], "little") for k in range(0, len(b), 32))

    def __eq__(self, other):
        return list(self) == list(other)


@dataclass
class PlacementVariables:
    subcircuitId: int
    variables: List[int]  # field elements (JSON: hex strings); a ScalarArray when read by the native loader


@dataclass
class Permutation:
    row: int
    col: int
    X: int
    Y: int


@dataclass
class Instance:
    a_pub_user: List[int]
    a_pub_block: List[int]
    a_pub_function: List[int]


# ---------------------------------------------------------------- iden3 .r1cs (R1csBinary, iotools/mod.rs:505-650)
@dataclass
class R1CS:
    n_wires: int
    n_constraints: int
    # constraints[row] = (A, B, C), each a list of (wire, coeff)
    constraints: list = field(default_factory=list)


def write_r1cs(path, r1cs: R1CS, n_pub_out=0, n_pub_in=0, n_prv_in=0):
    fs = 32
    header = struct.pack("<I", fs) + R_MOD.to_bytes(fs, "little") + struct.pack("<IIIIQI", r1cs.n_wires, n_pub_out, n_pub_in, n_prv_in,
                                                                              r1cs.n_wires, r1cs.n_constraints)
    body = bytearray()
    for abc in r1cs.constraints:
        for lc in abc:
            body += struct.pack("<I", len(lc))
            for wire, coeff in lc:
                body += struct.pack("<I", wire) + (coeff % R_MOD).to_bytes(fs, "little")
    wire_map = b"".join(struct.pack("<Q", i) for i in range(r1cs.n_wires))
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        for typ, sec in ((1, header), (2, bytes(body)), (3, wire_map)):
            f.write(struct.pack("<IQ", typ, len(sec)) + sec)


def read_r1cs(path) -> R1CS:
    data = open(path, "rb").read()
    if data[:4] != b"r1cs":
        raise ValueError("invalid R1CS magic")
    version, nsec = struct.unpack_from("<II", data, 4)
    if version != 1:
        raise ValueError(f"unsupported R1CS version {version}")
    off = 12
    secs = {}
    for _ in range(nsec):
        typ, size = struct.unpack_from("<IQ", data, off)
        off += 12
        if off + size > len(data):
            raise ValueError("R1CS section extends past end of file")
        secs.setdefault(typ, (off, size))
        off += size
    if 1 not in secs or 2 not in secs:
        raise ValueError("missing R1CS header or constraints section")
    h, _ = secs[1]
    (fs,) = struct.unpack_from("<I", data, h)
    if fs == 0 or fs % 8:
        raise ValueError(f"invalid R1CS field size {fs}")
    n_wires, _po, _pi, _pr, _nl, n_cons = struct.unpack_from("<IIIIQI", data, h + 4 + fs)
    c, csize = secs[2]
    end = c + csize
    out = R1CS(n_wires, n_cons)
    for _row in range(n_cons):
        abc = []
        for _m in range(3):
            (cnt,) = struct.unpack_from("<I", data, c)
            c += 4
            lc = []
            for _ in range(cnt):
                (wire,) = struct.unpack_from("<I", data, c)
                if wire >= n_wires:
                    raise ValueError(f"R1CS wire index {wire} exceeds nWires {n_wires}")
                lc.append((wire, int.from_bytes(data[c + 4:c + 4 + fs], "little") % R_MOD))
                c += 4 + fs
            abc.append(lc)
        out.constraints.append(tuple(abc))
    if c != end:
        raise ValueError(f"R1CS constraints section has {end - c} trailing bytes")
    return out


# ---------------------------------------------------------------- JSON artefacts
def _dump(path, obj):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        json.dump(obj, f)


def write_library(qap_path, params: SetupParams, infos, r1cs_list):
    """setupParams.json, subcircuitInfo.json, globalWireList.json and r1cs/subcircuit{i}.r1cs."""
    _dump(os.path.join(qap_path, "setupParams.json"), asdict(params))
    _dump(os.path.join(qap_path, "subcircuitInfo.json"), [asdict(s) for s in infos])
    gwl = [[-1, -1] for _ in range(params.m_D)]
    for s in infos:
        for local, g in enumerate(s.flattenMap):
            gwl[g] = [s.id, local]
    _dump(os.path.join(qap_path, "globalWireList.json"), gwl)
    os.makedirs(os.path.join(qap_path, "r1cs"), exist_ok=True)
    for s, r in zip(infos, r1cs_list):
        write_r1cs(os.path.join(qap_path, "r1cs", f"subcircuit{s.id}.r1cs"), r, n_pub_out=s.Out_idx[1], n_pub_in=s.In_idx[1])


def read_library_meta(qap_path):
    """setupParams.json + subcircuitInfo.json (+ the globalWireList consistency check) without the .r1cs binaries, which
    qap.library_csr_from_files parses natively."""
    params = SetupParams(**json.load(open(os.path.join(qap_path, "setupParams.json"))))
    infos = [SubcircuitInfo(**{k: d[k] for k in ("id", "name", "Nwires", "Nconsts", "Out_idx", "In_idx", "flattenMap")})
             for d in json.load(open(os.path.join(qap_path, "subcircuitInfo.json")))]
    gwl = json.load(open(os.path.join(qap_path, "globalWireList.json")))
    for s in infos:
        for local, g in enumerate(s.flattenMap):
            if gwl[g] != [s.id, local]:
                raise ValueError("GlobalWireList is not the inverse of flattenMap.")
    return params, infos


def read_packed_library(path):
    """The circuit library in the packed form tests/golden/gen_real_library.py writes (setupParams, subcircuitInfo, the
    .r1cs matrices as CSR row lengths / wire indices / indices into one coefficient table; lzma-compressed JSON):
    -> (SetupParams, [SubcircuitInfo], [R1CS]) like read_library."""
    import lzma

    with lzma.open(path, "rt") as f:
        doc = json.load(f)
    params = SetupParams(**doc["setupParams"])
    infos = [SubcircuitInfo(**d) for d in doc["subcircuitInfo"]]
    coeffs = [int(c, 16) for c in doc["coeffs"]]
    r1cs = []
    for s, d in zip(infos, doc["r1cs"]):
        if d["id"] != s.id or d["n_wires"] != s.Nwires or d["n_constraints"] != s.Nconsts or len(d["lens"]) != 3 * s.Nconsts:
            raise ValueError(f"packed library: shape mismatch for subcircuit {s.id}")
        cons, pos = [], 0
        wires, cidx, lens = d["wires"], d["coeff_idx"], d["lens"]
        for row in range(s.Nconsts):
            abc = []
            for m in range(3):
                k = lens[3 * row + m]
                abc.append([(wires[pos + t], coeffs[cidx[pos + t]]) for t in range(k)])
                pos += k
            cons.append(tuple(abc))
        if pos != len(wires) or any(w >= s.Nwires for w in wires):
            raise ValueError(f"packed library: corrupt CSR for subcircuit {s.id}")
        r1cs.append(R1CS(s.Nwires, s.Nconsts, cons))
    return params, infos, r1cs


def read_library(qap_path):
    params = SetupParams(**json.load(open(os.path.join(qap_path, "setupParams.json"))))
    infos = [SubcircuitInfo(**{k: d[k] for k in ("id", "name", "Nwires", "Nconsts", "Out_idx", "In_idx", "flattenMap")})
             for d in json.load(open(os.path.join(qap_path, "subcircuitInfo.json")))]
    gwl = json.load(open(os.path.join(qap_path, "globalWireList.json")))
    for s in infos:  # "GlobalWireList is not the inverse of flattenMap." (trusted-setup/src/main.rs:150-156)
        for local, g in enumerate(s.flattenMap):
            if gwl[g] != [s.id, local]:
                raise ValueError("GlobalWireList is not the inverse of flattenMap.")
    r1cs = []
    for s in infos:
        r = read_r1cs(os.path.join(qap_path, "r1cs", f"subcircuit{s.id}.r1cs"))
        if r.n_wires != s.Nwires or r.n_constraints != s.Nconsts:
            raise ValueError(f"R1CS shape mismatch for subcircuit {s.id}")
        if params.n < s.Nconsts:
            raise ValueError("n is smaller than the actual number of constraints.")
        r1cs.append(r)
    return params, infos, r1cs


def write_synthesizer_output(path, placements, permutation, instance: Instance):
    _dump(os.path.join(path, "placementVariables.json"), [{"subcircuitId": p.subcircuitId, "variables": [to_hex(v) for v in p.variables]} for p in placements])
    _dump(os.path.join(path, "permutation.json"), [asdict(p) for p in permutation])
    _dump(os.path.join(path, "instance.json"), {k: [to_hex(v) for v in getattr(instance, k)] for k in ("a_pub_user", "a_pub_block", "a_pub_function")})


def read_placement_variables_native(path, infos):
    """placementVariables.json through the library's host-side loader (tkm_host_parse_hex_scalars): the file is scanned
    once in native code, the values never become Python integers.  `infos` gives the wire count of every subcircuit."""
    import ctypes
    import re

    import numpy as np

    from .. import ffi

    data = open(os.path.join(path, "placementVariables.json"), "rb").read()
    ids = [int(m) for m in re.findall(rb'"subcircuitId"\s*:\s*(\d+)', data)]
    # per-placement counts before the flat stream is split (the reference checks each placement: iotools/mod.rs:505-520):
    # the hex strings of every "variables" array are counted with C-speed byte scans
    segs = data.split(b'"variables"')[1:]
    if len(segs) != len(ids):
        raise ValueError("Corrupted placement variables.")
    for i, seg in zip(ids, segs):
        end = seg.find(b']')
        if end < 0 or seg.count(b'"0x', 0, end) + seg.count(b'"0X', 0, end) != infos[i].Nwires:
            raise ValueError("Corrupted placement variables.")
    total = sum(infos[i].Nwires for i in ids)
    vals = np.empty((total, 4), dtype=np.uint64)
    count = ctypes.c_size_t()
    ffi.check(ffi.load().tkm_host_parse_hex_scalars(data, len(data), vals.ctypes.data_as(ctypes.c_void_p), total, ctypes.byref(count)))
    if count.value != total:
        raise ValueError("Corrupted placement variables.")
    out, off = [], 0
    for i in ids:
        k = infos[i].Nwires
        out.append(PlacementVariables(i, ScalarArray(vals[off:off + k])))
        off += k
    return out


def read_synthesizer_output(path, infos=None):
    """permutation.json, instance.json and placementVariables.json; with `infos` the 40 MB of hex strings go through the
    native loader."""
    if infos is not None:
        placements = read_placement_variables_native(path, infos)
        permutation = [Permutation(d["row"], d["col"], d["X"], d["Y"]) for d in json.load(open(os.path.join(path, "permutation.json")))]
        d = json.load(open(os.path.join(path, "instance.json")))
        return placements, permutation, Instance(*[[from_hex(v) for v in d[k]] for k in ("a_pub_user", "a_pub_block", "a_pub_function")])
    return _read_synthesizer_output_python(path)


def _read_synthesizer_output_python(path):
    # int(v, 16) accepts the 0x prefix; values above r (never produced by the synthesizer) are reduced like from_hex does
    placements = []
    for d in json.load(open(os.path.join(path, "placementVariables.json"))):
        vals = [int(v, 16) for v in d["variables"]]
        if vals and max(vals) >= R_MOD:
            vals = [v % R_MOD for v in vals]
        placements.append(PlacementVariables(d["subcircuitId"], vals))
    permutation = [Permutation(d["row"], d["col"], d["X"], d["Y"]) for d in json.load(open(os.path.join(path, "permutation.json")))]
    d = json.load(open(os.path.join(path, "instance.json")))
    inst = Instance(*[[from_hex(v) for v in d[k]] for k in ("a_pub_user", "a_pub_block", "a_pub_function")])
    return placements, permutation, inst


# ---------------------------------------------------------------- Solidity-split proof / preprocess JSON
PROOF_G1_ORDER = ["U", "V", "W", "O_mid", "O_prv", "Q_AX", "Q_AY", "Q_CX", "Q_CY", "Pi_X", "Pi_Y", "B", "R", "M_Y", "M_X", "N_Y", "N_X",
                  "O_pub_free", "A_free"]  # Proof::convert_format_for_solidity_verifier (prove/src/lib.rs:453-512)
PROOF_SCALAR_ORDER = ["R_eval", "R_omegaX_eval", "R_omegaX_omegaY_eval", "V_eval"]
PREPROCESS_G1_ORDER = ["s0", "s1", "O_pub_fix"]  # preprocess/src/lib.rs:84-106


def _split_fq(v):
    b = (v % Q_MOD).to_bytes(48, "big")
    return "0x" + b[:16].hex(), "0x" + b[16:].hex()


def _format_points(points, order):
    p1, p2 = [], []
    for name in order:
        pt = points[name]
        x, y = (0, 0) if pt is None else pt
        for c in (x, y):
            a, b = _split_fq(c)
            p1.append(a)
            p2.append(b)
    return p1, p2


def _recover_points(p1, p2, order):
    out = {}
    for i, name in enumerate(order):
        x = int(p1[2 * i][2:] + p2[2 * i][2:], 16)
        y = int(p1[2 * i + 1][2:] + p2[2 * i + 1][2:], 16)
        out[name] = None if x == 0 and y == 0 else (x, y)
    return out


def format_proof(points, scalars):
    """FormattedProof: 19 G1 points split 16+32 bytes big-endian per coordinate, then 4 scalars in part2."""
    p1, p2 = _format_points(points, PROOF_G1_ORDER)
    for name in PROOF_SCALAR_ORDER:
        p2.append("0x" + (scalars[name] % R_MOD).to_bytes(32, "big").hex())
    return {"proof_entries_part1": p1, "proof_entries_part2": p2}


def recover_proof(fmt):
    p1, p2 = fmt["proof_entries_part1"], fmt["proof_entries_part2"]
    n = len(PROOF_G1_ORDER)
    assert len(p1) == 2 * n and len(p2) == 2 * n + len(PROOF_SCALAR_ORDER)
    points = _recover_points(p1, p2, PROOF_G1_ORDER)
    scalars = {name: from_hex(p2[2 * n + i]) for i, name in enumerate(PROOF_SCALAR_ORDER)}
    return points, scalars


def format_preprocess(points):
    p1, p2 = _format_points(points, PREPROCESS_G1_ORDER)
    return {"preprocess_entries_part1": p1, "preprocess_entries_part2": p2}


def recover_preprocess(fmt):
    return _recover_points(fmt["preprocess_entries_part1"], fmt["preprocess_entries_part2"], PREPROCESS_G1_ORDER)
