"""CRS file formats of the reference.

Two containers hold `combined_sigma`:

* **TZBWASM1** -- the flat binary container of the browser prover (packages/backend-wasm/src/artifacts/binary/
  binary-format.ts:1-110, binary-artifact-file.ts:38-118; section list: src/artifacts/specs/prover-crs.v1.json).  64-byte
  header, file-kind table, version table, one self digest (SHA-256 of the file with the digest bytes zeroed), 96-byte
  section entries, 8-aligned section data.  Points are `ffjs-g1-affine-96` / `ffjs-g2-affine-192`: x || y, every Fq
  coordinate 48 bytes little-endian in MONTGOMERY form (R = 2^384) -- byte for byte the layout libtokamak_b200 keeps on the
  device, so a section is uploaded with tkm_crs_upload_mont without any conversion.
* **rkyv 0.7 archive** of `SigmaRkyv` -- what the native prover mmaps (libs/src/iotools/mod.rs:1701-1783,
  prove/src/sigma_source.rs:17-37): root object at the end of the file, `ArchivedVec` = 32-bit relative pointer + 32-bit
  length, points as canonical little-endian `[u8; 48]` pairs (BaseField::to_bytes_le).  The archived structs are
  `repr(Rust)`, so their field order is the compiler's: the reader accepts the declaration order and the
  alignment-sorted order rustc produces (4-aligned vectors first, byte arrays after) and tells them apart by checking
  that every vector lands inside the file and the fixed points decode on the curve.  No real archive exists in the
  reference tree, so this layout is restated from the derive rules, not pinned against a file the reference wrote.

Everything here is host-side codec work; the tables go to the device through `backend.table_from_points` /
`table_from_mont_bytes`.
"""
import hashlib
import struct

import numpy as np

from .fr import Q_MOD

MAGIC = b"TZBWASM1"
HEADER_BYTES, KIND_TABLE_BYTES, VERSION_TABLE_BYTES, DIGEST_ENTRY_BYTES, SECTION_ENTRY_BYTES, LABEL_BYTES = 64, 8, 72, 40, 96, 40
KIND_PROVER_CRS = 6
ENC_G1, ENC_G2 = 3, 4
TYPE_CRS_G1, TYPE_CRS_G2 = 11, 12
R384 = 1 << 384
R384_INV = pow(R384, -1, Q_MOD)

# prover-crs.v1.json: section labels in file order; the fixed points of "sigma.g1" / "sigma.g2"
G1_FIXED_NAMES = ["G", "sigma1.x", "sigma1.y", "sigma1.delta", "sigma1.eta", "lagrangeKL"]
G2_FIXED_NAMES = ["H", "sigma2.alpha", "sigma2.alpha2", "sigma2.alpha3", "sigma2.alpha4", "sigma2.gamma", "sigma2.delta", "sigma2.eta", "sigma2.x", "sigma2.y"]
G1_TABLE_LABELS = ["sigma1.xy-powers", "sigma1.gamma-inv-o-inst", "sigma1.eta-inv-li-o-inter-alpha4-kj", "sigma1.delta-inv-li-o-prv",
                   "sigma1.delta-inv-alphak-xh-tx", "sigma1.delta-inv-alpha4-xj-tx", "sigma1.delta-inv-alphak-yi-ty"]


def _align8(v):
    return (v + 7) & ~7


# ------------------------------------------------------------------------------------------ point codecs
def fq_to_mont_bytes(v):
    return (v * R384 % Q_MOD).to_bytes(48, "little")


def fq_from_mont_bytes(b):
    return int.from_bytes(b, "little") * R384_INV % Q_MOD


def g1_to_ffjs(pt):
    """(x, y) | None -> 96 bytes, Montgomery little-endian; the identity is all zero."""
    return bytes(96) if pt is None else fq_to_mont_bytes(pt[0]) + fq_to_mont_bytes(pt[1])


def g1_from_ffjs(b):
    b = bytes(b)
    if b == bytes(96):
        return None
    return fq_from_mont_bytes(b[:48]), fq_from_mont_bytes(b[48:])


def g2_to_ffjs(pt):
    (x0, x1), (y0, y1) = pt
    return b"".join(fq_to_mont_bytes(v) for v in (x0, x1, y0, y1))


def g2_from_ffjs(b):
    v = [fq_from_mont_bytes(bytes(b[48 * i:48 * i + 48])) for i in range(4)]
    return (v[0], v[1]), (v[2], v[3])


def g1_to_canonical(pt):
    return bytes(96) if pt is None else pt[0].to_bytes(48, "little") + pt[1].to_bytes(48, "little")


def g1_from_canonical(b):
    b = bytes(b)
    x, y = int.from_bytes(b[:48], "little"), int.from_bytes(b[48:], "little")
    return None if x == 0 and y == 0 else (x, y)


def g2_to_canonical(pt):
    (x0, x1), (y0, y1) = pt
    return b"".join(v.to_bytes(48, "little") for v in (x0, x1, y0, y1))


def g2_from_canonical(b):
    v = [int.from_bytes(bytes(b[48 * i:48 * i + 48]), "little") for i in range(4)]
    return (v[0], v[1]), (v[2], v[3])


def canonical_points_to_mont(points):
    """(n, 12) u64 canonical -> (n, 12) u64 Montgomery on the host (Python integers: for small tables and CPU tests; the GPU
    backend converts on the device)."""
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
    out = np.empty_like(pts)
    for i in range(pts.shape[0]):
        out[i] = np.frombuffer(g1_to_ffjs(g1_from_canonical(pts[i].tobytes())), dtype=np.uint64)
    return out


def mont_points_to_canonical(points):
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 12)
    out = np.empty_like(pts)
    for i in range(pts.shape[0]):
        out[i] = np.frombuffer(g1_to_canonical(g1_from_ffjs(pts[i].tobytes())), dtype=np.uint64)
    return out


def g1_on_curve(pt):
    return pt is None or (pt[0] < Q_MOD and pt[1] < Q_MOD and (pt[1] * pt[1] - pt[0] * pt[0] * pt[0] - 4) % Q_MOD == 0)


# ------------------------------------------------------------------------------------------ TZBWASM1 container
def write_tzbwasm(path, kind, source_package_version, sections):
    """createBinaryArtifactFile (binary-artifact-file.ts:38-118).  sections: dicts with type, encoding, label, element_count,
    element_bytes and data (bytes-like; numpy arrays are written without copying)."""
    ver = source_package_version.encode()
    if not ver or source_package_version.strip() != source_package_version or len(ver) > 64:
        raise ValueError("Binary artifact sourcePackageVersion must be a non-empty trimmed string.")
    kind_off = HEADER_BYTES
    ver_off = kind_off + KIND_TABLE_BYTES
    dig_off = ver_off + VERSION_TABLE_BYTES
    sec_off = _align8(dig_off + DIGEST_ENTRY_BYTES)
    data_off = _align8(sec_off + len(sections) * SECTION_ENTRY_BYTES)
    offs, off = [], data_off
    for s in sections:
        nbytes = memoryview(s["data"]).nbytes
        if s["element_count"] * s["element_bytes"] != nbytes:
            raise ValueError(f"Section '{s['label']}' byte length does not match its element count.")
        if len(s["label"].encode()) > LABEL_BYTES:
            raise ValueError(f"Binary section label is longer than {LABEL_BYTES} bytes: {s['label']}.")
        off = _align8(off)
        offs.append(off)
        off += nbytes
    total = _align8(max([data_off] + [o + memoryview(s["data"]).nbytes for o, s in zip(offs, sections)]))
    if total >= 1 << 32:
        raise ValueError("binary artifact exceeds the format's 32-bit offsets")
    head = bytearray(data_off)
    head[0:8] = MAGIC
    struct.pack_into("<H", head, 8, 1)        # format version; bytes 10-11 stay zero
    struct.pack_into("<I", head, 12, total)   # byteLength
    struct.pack_into("<IIIIIIIII", head, 16, kind_off, KIND_TABLE_BYTES, ver_off, VERSION_TABLE_BYTES, dig_off, DIGEST_ENTRY_BYTES, sec_off,
                     len(sections) * SECTION_ENTRY_BYTES, data_off)
    struct.pack_into("<HHII", head, 52, len(sections), 1, 0, 0)
    struct.pack_into("<HHI", head, kind_off, kind, 0, 0)
    struct.pack_into("<HHI", head, ver_off, 1, len(ver), 0)
    head[ver_off + 8:ver_off + 8 + len(ver)] = ver
    struct.pack_into("<HHI", head, dig_off, 1, 0xFFFF, 0)
    for i, (s, o) in enumerate(zip(sections, offs)):
        e = sec_off + i * SECTION_ENTRY_BYTES
        struct.pack_into("<HHIIIIHH", head, e, s["type"], s["encoding"], s.get("flags", 0), o, memoryview(s["data"]).nbytes, s["element_count"],
                         s["element_bytes"], 0)
        lab = s["label"].encode()
        head[e + 56:e + 56 + len(lab)] = lab
    # self digest: SHA-256 over the whole file with the 32 digest bytes zero
    h = hashlib.sha256()
    h.update(bytes(head))
    pos = data_off
    chunks = []
    for s, o in zip(sections, offs):
        if o > pos:
            chunks.append(bytes(o - pos))
        chunks.append(memoryview(s["data"]).cast("B"))
        pos = o + memoryview(s["data"]).nbytes
    if total > pos:
        chunks.append(bytes(total - pos))
    for c in chunks:
        h.update(c)
    head[dig_off + 8:dig_off + 40] = h.digest()
    with open(path, "wb") as f:
        f.write(head)
        for c in chunks:
            f.write(c)


def read_tzbwasm(path, verify_digest=True):
    """decodeBinaryArtifactFile (:120-181): {"kind", "format_version", "source_package_version", "sections": [...]}; section
    data are zero-copy views of a read-only memory map."""
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if mm.shape[0] < HEADER_BYTES:
        raise ValueError("Binary artifact is shorter than the fixed header.")
    if bytes(mm[0:8]) != MAGIC:
        raise ValueError("not a TZBWASM1 binary artifact")
    hdr = bytes(mm[:HEADER_BYTES])
    fmt = struct.unpack_from("<H", hdr, 8)[0]
    byte_length = struct.unpack_from("<I", hdr, 12)[0]
    kind_off, _, ver_off, ver_len, dig_off, _, sec_off, _, _ = struct.unpack_from("<IIIIIIIII", hdr, 16)
    n_sec, n_dig = struct.unpack_from("<HH", hdr, 52)
    size = mm.shape[0]

    def rng(off, length, what):
        if off > size or length > size - off:
            raise ValueError(f"{what} extends outside the binary artifact input.")

    if byte_length != size:
        raise ValueError("binary artifact byteLength field does not match the file size")
    rng(kind_off, 2, "binary artifact file-kind table")
    kind = struct.unpack_from("<H", bytes(mm[kind_off:kind_off + 2]))[0]
    rng(ver_off, ver_len, "binary artifact version table")
    if ver_len < 8:
        raise ValueError("Binary artifact version table is too short to read.")
    vlen = struct.unpack_from("<H", bytes(mm[ver_off + 2:ver_off + 4]))[0]
    version = bytes(mm[ver_off + 8:ver_off + 8 + vlen]).decode()
    if n_dig != 1:
        raise ValueError("Binary artifact must contain exactly one self digest.")
    rng(dig_off, DIGEST_ENTRY_BYTES, "binary artifact self digest")
    dt, di = struct.unpack_from("<HH", bytes(mm[dig_off:dig_off + 4]))
    if dt != 1 or di != 0xFFFF:
        raise ValueError("Binary artifact digest table must contain only the self digest.")
    digest = bytes(mm[dig_off + 8:dig_off + 40])
    if verify_digest:
        h = hashlib.sha256()
        h.update(mm[:dig_off + 8])
        h.update(bytes(32))
        step = 1 << 26
        for o in range(dig_off + 40, size, step):
            h.update(mm[o:min(size, o + step)])
        if h.digest() != digest:
            raise ValueError("binary artifact self digest mismatch")
    sections = []
    for i in range(n_sec):
        e = sec_off + i * SECTION_ENTRY_BYTES
        rng(e, SECTION_ENTRY_BYTES, f"binary artifact section table entry {i}")
        ent = bytes(mm[e:e + SECTION_ENTRY_BYTES])
        typ, enc, flags, off, length, count, ebytes, _ = struct.unpack_from("<HHIIIIHH", ent, 0)
        label = ent[56:56 + LABEL_BYTES].split(b"\0")[0].decode()
        rng(off, length, f"binary artifact section '{label}'")
        if count * ebytes != length:
            raise ValueError(f"Section '{label}' byte length does not match its element count.")
        sections.append({"type": typ, "encoding": enc, "flags": flags, "label": label, "element_count": count, "element_bytes": ebytes,
                         "byte_offset": off, "data": mm[off:off + length]})
    return {"kind": kind, "format_version": fmt, "source_package_version": version, "self_digest": digest, "sections": sections}


# ------------------------------------------------------------------------------------------ Sigma <-> prover_crs (TZBWASM1)
def _table_mont_bytes(backend, table):
    """rows*cols x 96 bytes in the ffjs (= device) layout."""
    if hasattr(table, "mont_bytes_host"):
        return table.mont_bytes_host()
    return canonical_points_to_mont(table.points_host())


def write_prover_crs(path, backend, sigma, source_package_version="tokamak-b200/0.1.0"):
    """combined sigma -> prover_crs artifact (specs/prover-crs.v1.json)."""
    s2 = sigma.sigma2
    g1_fixed = [sigma.G, sigma.x, sigma.y, sigma.delta, sigma.eta, sigma.lagrange_KL]
    g2_fixed = [sigma.H, s2.alpha, s2.alpha2, s2.alpha3, s2.alpha4, s2.gamma, s2.delta, s2.eta, s2.x, s2.y]

    def g1_section(label, data, count):
        return {"type": TYPE_CRS_G1, "encoding": ENC_G1, "label": label, "element_count": count, "element_bytes": 96, "data": data}

    sections = [g1_section("sigma.g1", b"".join(g1_to_ffjs(p) for p in g1_fixed), 6)]
    for label, t in (("sigma1.xy-powers", sigma.xy_powers), ("sigma1.gamma-inv-o-inst", sigma.gamma_inv_o_inst),
                     ("sigma1.eta-inv-li-o-inter-alpha4-kj", sigma.eta_inv_li_o_inter_alpha4_kj), ("sigma1.delta-inv-li-o-prv", sigma.delta_inv_li_o_prv)):
        sections.append(g1_section(label, np.ascontiguousarray(_table_mont_bytes(backend, t)).view(np.uint8).reshape(-1), t.rows * t.cols))
    xh = [p for row in sigma.delta_inv_alphak_xh_tx for p in row]
    yi = [p for row in sigma.delta_inv_alphak_yi_ty for p in row]
    sections.append(g1_section("sigma1.delta-inv-alphak-xh-tx", b"".join(g1_to_ffjs(p) for p in xh), len(xh)))
    sections.append(g1_section("sigma1.delta-inv-alpha4-xj-tx", b"".join(g1_to_ffjs(p) for p in sigma.delta_inv_alpha4_xj_tx), len(sigma.delta_inv_alpha4_xj_tx)))
    sections.append(g1_section("sigma1.delta-inv-alphak-yi-ty", b"".join(g1_to_ffjs(p) for p in yi), len(yi)))
    sections.append({"type": TYPE_CRS_G2, "encoding": ENC_G2, "label": "sigma.g2", "element_count": 10, "element_bytes": 192,
                     "data": b"".join(g2_to_ffjs(p) for p in g2_fixed)})
    write_tzbwasm(path, KIND_PROVER_CRS, source_package_version, sections)


def _upload_mont(backend, data, rows, cols):
    if hasattr(backend, "table_from_mont_bytes"):
        return backend.table_from_mont_bytes(data, rows, cols)  # device layout already: no conversion
    return backend.table_from_points(mont_points_to_canonical(np.frombuffer(bytes(data), dtype=np.uint64)), rows, cols)


def read_prover_crs(path, backend, params, verify_digest=True):
    """prover_crs artifact -> Sigma with its four large tables resident on `backend` (shapes from the setup parameters:
    xy_powers [max(2n, 2 m_I)][2 s_max], gamma_inv_o_inst [l][1], eta_inv_li_o_inter_alpha4_kj [m_I][s_max],
    delta_inv_li_o_prv [m_D - l_D][s_max]; libs/src/group_structures/mod.rs:361-394)."""
    from .setup import Sigma, Sigma2

    art = read_tzbwasm(path, verify_digest)
    if art["kind"] != KIND_PROVER_CRS:
        raise ValueError(f"not a prover_crs artifact (file kind {art['kind']})")
    sec = {}
    for s in art["sections"]:
        want = (TYPE_CRS_G2, ENC_G2, 192) if s["label"] == "sigma.g2" else (TYPE_CRS_G1, ENC_G1, 96)
        if (s["type"], s["encoding"], s["element_bytes"]) != want:
            raise ValueError(f"section '{s['label']}' has an unexpected type/encoding")
        sec[s["label"]] = s
    for label in ["sigma.g1", "sigma.g2"] + G1_TABLE_LABELS:
        if label not in sec:
            raise ValueError(f"Missing binary artifact section: {label}.")
    p = params
    m_i = p.l_D - p.l
    shapes = {"sigma1.xy-powers": (max(2 * p.n, 2 * m_i), 2 * p.s_max), "sigma1.gamma-inv-o-inst": (p.l, 1),
              "sigma1.eta-inv-li-o-inter-alpha4-kj": (m_i, p.s_max), "sigma1.delta-inv-li-o-prv": (p.m_D - p.l_D, p.s_max),
              "sigma1.delta-inv-alphak-xh-tx": (3, 3), "sigma1.delta-inv-alpha4-xj-tx": (2, 1), "sigma1.delta-inv-alphak-yi-ty": (4, 3)}
    for label, (r, c) in shapes.items():
        if sec[label]["element_count"] != r * c:
            raise ValueError(f"section '{label}' holds {sec[label]['element_count']} points, the setup parameters need {r} x {c}")
    if sec["sigma.g1"]["element_count"] != 6 or sec["sigma.g2"]["element_count"] != 10:
        raise ValueError("fixed-point sections have the wrong element count")
    sg = Sigma()
    g1 = [g1_from_ffjs(sec["sigma.g1"]["data"][96 * i:96 * i + 96]) for i in range(6)]
    for pt in g1:
        if not g1_on_curve(pt):
            raise ValueError("a fixed G1 point of the CRS is not on the curve")
    sg.G, sg.x, sg.y, sg.delta, sg.eta, sg.lagrange_KL = g1
    g2 = [g2_from_ffjs(sec["sigma.g2"]["data"][192 * i:192 * i + 192]) for i in range(10)]
    sg.H = g2[0]
    sg.sigma2 = Sigma2(*g2[1:])
    tab = lambda label: _upload_mont(backend, sec[label]["data"], *shapes[label])
    sg.xy_powers, sg.gamma_inv_o_inst = tab("sigma1.xy-powers"), tab("sigma1.gamma-inv-o-inst")
    sg.eta_inv_li_o_inter_alpha4_kj, sg.delta_inv_li_o_prv = tab("sigma1.eta-inv-li-o-inter-alpha4-kj"), tab("sigma1.delta-inv-li-o-prv")
    small = lambda label, n: [g1_from_ffjs(sec[label]["data"][96 * i:96 * i + 96]) for i in range(n)]
    xh, xj, yi = small("sigma1.delta-inv-alphak-xh-tx", 9), small("sigma1.delta-inv-alpha4-xj-tx", 2), small("sigma1.delta-inv-alphak-yi-ty", 12)
    sg.delta_inv_alphak_xh_tx = [xh[3 * k:3 * k + 3] for k in range(3)]
    sg.delta_inv_alpha4_xj_tx = xj
    sg.delta_inv_alphak_yi_ty = [yi[3 * k:3 * k + 3] for k in range(4)]
    return sg


# ------------------------------------------------------------------------------------------ rkyv 0.7 SigmaRkyv archive
# Archived sizes: G1SerdeRkyv 96 (align 1), G2SerdeRkyv 192 (align 1), ArchivedVec 8 (align 4: i32 relative offset from the
# pointer's own position, u32 length).  Sigma2Rkyv = 9 x 192 = 1728 (align 1).
_S1_FIELDS = [("xy_powers", "vec"), ("x", "g1"), ("y", "g1"), ("delta", "g1"), ("eta", "g1"), ("gamma_inv_o_inst", "vec"),
              ("eta_inv_li_o_inter_alpha4_kj", "vecvec"), ("delta_inv_li_o_prv", "vecvec"), ("delta_inv_alphak_xh_tx", "vecvec"),
              ("delta_inv_alpha4_xj_tx", "vec"), ("delta_inv_alphak_yi_ty", "vecvec")]
_SG_FIELDS = [("G", "g1"), ("H", "g2"), ("sigma_1", "s1"), ("sigma_2", "s2"), ("lagrange_KL", "g1")]
_SIZES = {"g1": 96, "g2": 192, "vec": 8, "vecvec": 8, "s2": 1728}
_ALIGN = {"g1": 1, "g2": 1, "vec": 4, "vecvec": 4, "s2": 1, "s1": 4}
LAYOUTS = ("aligned-first", "declaration")  # rustc's current reordering, and plain declaration order


def _struct_layout(fields, order, sizes):
    """[(name, kind, offset)], size: fields placed in `order` with natural alignment, size rounded to the struct alignment."""
    seq = list(fields)
    if order == "aligned-first":  # stable sort by descending alignment (rustc: larger alignment groups first)
        seq = sorted(seq, key=lambda f: -_ALIGN[f[1]])
    off, out, amax = 0, [], 1
    for name, kind in seq:
        a = _ALIGN[kind]
        amax = max(amax, a)
        off = (off + a - 1) // a * a
        out.append((name, kind, off))
        off += sizes[kind]
    return out, (off + amax - 1) // amax * amax


def _layouts(order):
    s1, s1_size = _struct_layout(_S1_FIELDS, order, _SIZES)
    sg, sg_size = _struct_layout(_SG_FIELDS, order, {**_SIZES, "s1": s1_size})
    return s1, s1_size, sg, sg_size


def write_sigma_rkyv(path, backend, sigma, layout="aligned-first"):
    """SigmaRkyv archive as rkyv 0.7's AllocSerializer lays it out: out-of-line vector data first (inner vectors before the
    table of their ArchivedVec headers), the root object last."""
    s1_l, s1_size, sg_l, sg_size = _layouts(layout)
    buf = bytearray()

    def put_points(arr_bytes):
        pos = len(buf)
        buf.extend(arr_bytes)
        return pos

    def table_bytes(t):
        return np.ascontiguousarray(t.points_host(), dtype=np.uint64).tobytes()

    def align(a):
        while len(buf) % a:
            buf.append(0)

    vec_pos = {}   # name -> (data position, length) for Vec<G1>
    vv_pos = {}    # name -> (header table position, outer length)
    for name, t in (("xy_powers", sigma.xy_powers), ("gamma_inv_o_inst", sigma.gamma_inv_o_inst)):
        vec_pos[name] = (put_points(table_bytes(t)), t.rows * t.cols)
    nested = {"eta_inv_li_o_inter_alpha4_kj": None, "delta_inv_li_o_prv": None}
    for name, t in (("eta_inv_li_o_inter_alpha4_kj", sigma.eta_inv_li_o_inter_alpha4_kj), ("delta_inv_li_o_prv", sigma.delta_inv_li_o_prv)):
        data = put_points(table_bytes(t))
        nested[name] = [(data + r * t.cols * 96, t.cols) for r in range(t.rows)]
    for name, rows in (("delta_inv_alphak_xh_tx", sigma.delta_inv_alphak_xh_tx), ("delta_inv_alphak_yi_ty", sigma.delta_inv_alphak_yi_ty)):
        nested[name] = []
        for row in rows:
            nested[name].append((put_points(b"".join(g1_to_canonical(p) for p in row)), len(row)))
    vec_pos["delta_inv_alpha4_xj_tx"] = (put_points(b"".join(g1_to_canonical(p) for p in sigma.delta_inv_alpha4_xj_tx)), len(sigma.delta_inv_alpha4_xj_tx))
    for name, inner in nested.items():
        align(4)
        pos = len(buf)
        for k, (data, n) in enumerate(inner):
            here = pos + 8 * k
            buf.extend(struct.pack("<iI", data - here, n))
        vv_pos[name] = (pos, len(inner))
    align(4)
    root = len(buf)
    buf.extend(bytes(sg_size))
    s2 = sigma.sigma2
    fixed = {"G": g1_to_canonical(sigma.G), "H": g2_to_canonical(sigma.H), "lagrange_KL": g1_to_canonical(sigma.lagrange_KL),
             "sigma_2": b"".join(g2_to_canonical(p) for p in (s2.alpha, s2.alpha2, s2.alpha3, s2.alpha4, s2.gamma, s2.delta, s2.eta, s2.x, s2.y))}
    s1_fixed = {"x": sigma.x, "y": sigma.y, "delta": sigma.delta, "eta": sigma.eta}
    for name, kind, off in sg_l:
        at = root + off
        if kind == "s1":
            for n1, k1, o1 in s1_l:
                a1 = at + o1
                if k1 == "g1":
                    buf[a1:a1 + 96] = g1_to_canonical(s1_fixed[n1])
                elif k1 == "vec":
                    data, n = vec_pos[n1]
                    buf[a1:a1 + 8] = struct.pack("<iI", data - a1, n)
                else:
                    pos, n = vv_pos[n1]
                    buf[a1:a1 + 8] = struct.pack("<iI", pos - a1, n)
        else:
            b = fixed[name]
            buf[at:at + len(b)] = b
    with open(path, "wb") as f:
        f.write(buf)


def read_sigma_rkyv(path, backend, params):
    """rkyv::archived_root::<SigmaRkyv> restated: the root object is the last size_of::<ArchivedSigmaRkyv>() bytes; vectors
    are followed through their relative pointers.  Returns (Sigma, layout name)."""
    from .setup import Sigma, Sigma2

    mm = np.memmap(path, dtype=np.uint8, mode="r")
    size = mm.shape[0]
    p = params
    m_i = p.l_D - p.l
    want = {"xy_powers": max(2 * p.n, 2 * m_i) * 2 * p.s_max, "gamma_inv_o_inst": p.l, "delta_inv_alpha4_xj_tx": 2}
    want_nested = {"eta_inv_li_o_inter_alpha4_kj": (m_i, p.s_max), "delta_inv_li_o_prv": (p.m_D - p.l_D, p.s_max),
                   "delta_inv_alphak_xh_tx": (3, 3), "delta_inv_alphak_yi_ty": (4, 3)}
    errors = []
    for layout in LAYOUTS:
        s1_l, s1_size, sg_l, sg_size = _layouts(layout)
        if size < sg_size:
            errors.append(f"{layout}: file shorter than the root object")
            continue
        root = size - sg_size
        try:
            def vec_at(a, elem):
                rel, n = struct.unpack_from("<iI", bytes(mm[a:a + 8]))
                tgt = a + rel
                if tgt < 0 or tgt + n * elem > root:
                    raise ValueError("vector outside the archive")
                return tgt, n

            sg = Sigma()
            top = {n_: root + o for n_, _, o in sg_l}
            sg.G = g1_from_canonical(mm[top["G"]:top["G"] + 96])
            sg.lagrange_KL = g1_from_canonical(mm[top["lagrange_KL"]:top["lagrange_KL"] + 96])
            if not (g1_on_curve(sg.G) and g1_on_curve(sg.lagrange_KL)) or sg.G is None:
                raise ValueError("fixed points are not on the curve")
            sg.H = g2_from_canonical(mm[top["H"]:top["H"] + 192])
            g2 = [g2_from_canonical(mm[top["sigma_2"] + 192 * i:top["sigma_2"] + 192 * i + 192]) for i in range(9)]
            sg.sigma2 = Sigma2(*g2)
            s1 = {n_: (k_, top["sigma_1"] + o) for n_, k_, o in s1_l}
            for nm in ("x", "y", "delta", "eta"):
                pt = g1_from_canonical(mm[s1[nm][1]:s1[nm][1] + 96])
                if not g1_on_curve(pt):
                    raise ValueError(f"sigma_1.{nm} is not on the curve")
                setattr(sg, nm, pt)
            tabs = {}
            for nm, cnt in want.items():
                tgt, n = vec_at(s1[nm][1], 96)
                if n != cnt:
                    raise ValueError(f"{nm} holds {n} points, the setup parameters need {cnt}")
                tabs[nm] = mm[tgt:tgt + n * 96]
            rows_of = {}
            for nm, (r, c) in want_nested.items():
                tgt, n = vec_at(s1[nm][1], 8)
                if n != r:
                    raise ValueError(f"{nm} holds {n} rows, the setup parameters need {r}")
                rows = []
                for k in range(n):
                    t2, n2 = vec_at(tgt + 8 * k, 96)
                    if n2 != c:
                        raise ValueError(f"{nm}[{k}] holds {n2} points, the setup parameters need {c}")
                    rows.append(mm[t2:t2 + n2 * 96])
                rows_of[nm] = rows
        except (ValueError, struct.error) as e:
            errors.append(f"{layout}: {e}")
            continue
        u64 = lambda v: np.frombuffer(np.ascontiguousarray(v).tobytes(), dtype=np.uint64)
        sg.xy_powers = backend.table_from_points(u64(tabs["xy_powers"]), max(2 * p.n, 2 * m_i), 2 * p.s_max)
        sg.gamma_inv_o_inst = backend.table_from_points(u64(tabs["gamma_inv_o_inst"]), p.l, 1)
        for nm in ("eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
            r, c = want_nested[nm]
            setattr(sg, nm, backend.table_from_points(u64(np.concatenate(rows_of[nm])), r, c))
        pts = lambda v, n: [g1_from_canonical(v[96 * i:96 * i + 96]) for i in range(n)]
        sg.delta_inv_alphak_xh_tx = [pts(rw, 3) for rw in rows_of["delta_inv_alphak_xh_tx"]]
        sg.delta_inv_alphak_yi_ty = [pts(rw, 3) for rw in rows_of["delta_inv_alphak_yi_ty"]]
        sg.delta_inv_alpha4_xj_tx = pts(tabs["delta_inv_alpha4_xj_tx"], 2)
        return sg, layout
    raise ValueError("Invalid sigma archive: " + "; ".join(errors))
