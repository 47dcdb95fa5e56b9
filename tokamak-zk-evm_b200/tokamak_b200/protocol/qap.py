"""Sparse R1CS products: the QAP mixture o_j(tau) for setup (host side like the reference's,
field_structures/mod.rs:73-151) and the data layouts the device kernels take for the prover's witness polynomials
(LibraryCSR + WitnessTable -> tkm_r1cs_uvw_polys; the literal per-placement host loop of iotools/mod.rs:1380-1608 lives with
the test oracle)."""
import numpy as np

from .fr import R_MOD, lagrange_bases_at


def o_evaled(params, infos, r1cs_list, tau):
    """o_vec[j] = alpha u_j(x) + alpha^2 v_j(x) + alpha^3 w_j(x) for every global wire j
    (from_r1cs_to_evaled_qap_mixture, libs/src/field_structures/mod.rs:73-151; scatter by flattenMap,
    setup/trusted-setup/src/main.rs:128-164).  u_j(X) = sum_rows A[row][j] K_row(X) over the n-th roots of unity."""
    lag = lagrange_bases_at(tau.x, params.n)
    a1, a2, a3 = tau.alpha % R_MOD, pow(tau.alpha, 2, R_MOD), pow(tau.alpha, 3, R_MOD)
    o = [0] * params.m_D
    for info, r in zip(infos, r1cs_list):
        local = [0] * info.Nwires
        for row, (a, b, c) in enumerate(r.constraints):
            lr = lag[row]
            for lc, scale in ((a, a1), (b, a2), (c, a3)):
                f = lr * scale % R_MOD
                for wire, coeff in lc:
                    local[wire] = (local[wire] + coeff * f) % R_MOD
        for loc, g in enumerate(info.flattenMap):
            if local[loc]:
                o[g] = local[loc]
    return o


class LibraryCSR:
    """The library's constraints as one concatenated CSR (per subcircuit and matrix), the layout tkm_r1cs_uvw_polys takes.
    Built once per library (the reference re-parses the .r1cs binaries on every prove, iotools/mod.rs:1322-1340)."""

    def __init__(self, r1cs_list):
        s_d = len(r1cs_list)
        self.n_rows = np.array([r.n_constraints for r in r1cs_list], dtype=np.uint32)
        self.rp_base = np.zeros(s_d * 3, dtype=np.uint64)
        row_ptr, wire, coeff = [], [], bytearray()
        for s, r in enumerate(r1cs_list):
            for m in range(3):
                self.rp_base[3 * s + m] = len(row_ptr)
                row_ptr.append(len(wire))
                for abc in r.constraints:
                    for wi, cf in abc[m]:
                        wire.append(wi)
                        coeff += cf.to_bytes(32, "little")
                    row_ptr.append(len(wire))
        self.row_ptr = np.array(row_ptr, dtype=np.uint32)
        self.wire = np.array(wire, dtype=np.uint32)
        self.coeff = np.frombuffer(bytes(coeff), dtype=np.uint64).reshape(-1, 4).copy()
        self.r1cs_list = r1cs_list


def library_csr_from_files(qap_path, params, infos):
    """LibraryCSR straight from r1cs/subcircuit{i}.r1cs through the library's host-side parser (tkm_host_parse_r1cs): no
    Python-level constraint lists are built.  Checks nWires / nConstraints against subcircuitInfo.json and n like
    SubcircuitR1CS::from_r1cs_with_mode (iotools/mod.rs:714-734)."""
    import ctypes
    import os

    from .. import ffi

    lib = ffi.load()
    s_d = len(infos)
    csr = LibraryCSR.__new__(LibraryCSR)
    csr.n_rows = np.zeros(s_d, dtype=np.uint32)
    csr.rp_base = np.zeros(s_d * 3, dtype=np.uint64)
    rps, wires, coeffs = [], [], []
    rp_off = ent_off = 0
    for s, info in enumerate(infos):
        data = open(os.path.join(qap_path, "r1cs", f"subcircuit{info.id}.r1cs"), "rb").read()
        nw, nc = ctypes.c_uint32(), ctypes.c_uint32()
        nnz = (ctypes.c_size_t * 3)()
        ffi.check(lib.tkm_host_parse_r1cs(data, len(data), ctypes.byref(nw), ctypes.byref(nc), nnz, None, None, None))
        if nw.value != info.Nwires or nc.value != info.Nconsts:
            raise ValueError(f"R1CS shape mismatch for subcircuit {info.id}")
        if params.n < nc.value:
            raise ValueError("n is smaller than the actual number of constraints.")
        total = nnz[0] + nnz[1] + nnz[2]
        rp = np.zeros(3 * (nc.value + 1), dtype=np.uint32)
        wi = np.zeros(max(total, 1), dtype=np.uint32)
        co = np.zeros((max(total, 1), 4), dtype=np.uint64)
        ffi.check(lib.tkm_host_parse_r1cs(data, len(data), ctypes.byref(nw), ctypes.byref(nc), nnz, rp.ctypes.data_as(ctypes.c_void_p),
                                          wi.ctypes.data_as(ctypes.c_void_p), co.ctypes.data_as(ctypes.c_void_p)))
        csr.n_rows[s] = nc.value
        for m in range(3):
            csr.rp_base[3 * s + m] = rp_off + m * (nc.value + 1)
        rps.append(rp + np.uint32(ent_off))
        wires.append(wi[:total])
        coeffs.append(co[:total])
        rp_off += rp.shape[0]
        ent_off += total
    csr.row_ptr = np.concatenate(rps)
    csr.wire = np.concatenate(wires) if ent_off else np.zeros(0, dtype=np.uint32)
    csr.coeff = np.concatenate(coeffs) if ent_off else np.zeros((0, 4), dtype=np.uint64)
    csr.r1cs_list = None
    return csr


class WitnessTable:
    """All placement variables as one (total, 4) uint64 array of canonical limbs + per-column offsets: the in-memory form
    of placementVariables.json the vectorised host code and the device kernels work on."""

    def __init__(self, params, placements, infos):
        if len(placements) > params.s_max:
            raise ValueError("placement_variables length exceeds s_max.")
        self.placements = placements
        self.sub_of_col = np.full(params.s_max, 0xFFFFFFFF, dtype=np.uint32)
        self.var_off = np.zeros(params.s_max, dtype=np.uint64)
        off = 0
        chunks = []
        for col, pl in enumerate(placements):
            if len(pl.variables) != infos[pl.subcircuitId].Nwires:
                raise ValueError("Corrupted placement variables.")
            self.sub_of_col[col] = pl.subcircuitId
            self.var_off[col] = off
            off += len(pl.variables)
            limbs = getattr(pl.variables, "limbs", None)  # ScalarArray from the native loader: already in device layout
            chunks.append(limbs if limbs is not None
                          else np.frombuffer(b"".join([v.to_bytes(32, "little") for v in pl.variables]), dtype=np.uint64).reshape(-1, 4))
        self.values = np.concatenate(chunks) if chunks else np.zeros((0, 4), dtype=np.uint64)  # one copy of the ~19 MB table
        self.fmap = [np.array(s.flattenMap, dtype=np.int64) for s in infos]

    def gather_indices(self, lo, hi, s_max):
        """Every (placement, wire) whose global wire index lies in [lo, hi): -> (table index (g - lo) * s_max + col, row of
        the value in self.values), both uint32.  The selection depends only on the subcircuit, so it is computed once per
        subcircuit and shifted per placement."""
        cache = self.__dict__.setdefault("_sel_cache", {})
        idx, rows = [], []
        for col, pl in enumerate(self.placements):
            key = (pl.subcircuitId, lo, hi, s_max)
            if key not in cache:
                fm = self.fmap[pl.subcircuitId]
                sel = np.nonzero((fm >= lo) & (fm < hi))[0]
                cache[key] = ((fm[sel] - lo) * s_max, sel)
            base_idx, sel = cache[key]
            idx.append(base_idx + col)
            rows.append(sel + int(self.var_off[col]))
        idx = np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)
        rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
        return idx.astype(np.uint32), rows.astype(np.uint32)

    def gather(self, lo, hi, s_max):
        """gather_indices with the values collected on the host: -> (table index, values)."""
        idx, rows = self.gather_indices(lo, hi, s_max)
        return idx, np.ascontiguousarray(self.values[rows])


def interface_evals_from_table(params, wt: WitnessTable):
    """gen_bXY over a WitnessTable (vectorised)."""
    s_max = params.s_max
    out = np.zeros(((params.l_D - params.l) * s_max, 4), dtype=np.uint64)
    idx, vals = wt.gather(params.l, params.l_D, s_max)
    out[idx] = vals
    return out


def permutation_evals(permutation, m_i, s_max, omega_m_i, omega_s_max):
    """Evaluation tables of s0, s1 (Permutation::to_poly, iotools/mod.rs:419-455): identity w_x^row, w_y^col except at the
    listed (row, col), which point to (X, Y)."""
    from .fr import powers

    xp, yp = powers(omega_m_i, m_i), powers(omega_s_max, s_max)
    mask = (1 << 64) - 1

    def limbs(v):
        return (v & mask, (v >> 64) & mask, (v >> 128) & mask, v >> 192)

    xl = np.array([limbs(v) for v in xp], dtype=np.uint64)
    yl = np.array([limbs(v) for v in yp], dtype=np.uint64)
    s0 = np.repeat(xl, s_max, axis=0)
    s1 = np.tile(yl, (m_i, 1))
    if len(permutation):
        e = np.array([(p.row, p.col, p.X, p.Y) for p in permutation], dtype=np.int64)
        idx = e[:, 0] * s_max + e[:, 1]
        # a later entry for the same (row, col) overrides an earlier one, like the reference's sequential loop
        _, last = np.unique(idx[::-1], return_index=True)
        keep = len(idx) - 1 - last
        s0[idx[keep]] = xl[e[keep, 2]]
        s1[idx[keep]] = yl[e[keep, 3]]
    return s0, s1
