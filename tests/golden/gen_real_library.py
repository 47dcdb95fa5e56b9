"""Packs the reference's checked-in circuit library (packages/frontend/qap-compiler/subcircuits/library: setupParams.json,
subcircuitInfo.json and the 14 iden3 .r1cs binaries) into tests/golden/real_library.json.xz.

The fixture carries the REAL constraint structure the prover's witness kernels see -- per subcircuit and matrix the CSR row
lengths, wire indices and coefficients (470 distinct values, stored once) -- in about 70 KB instead of 3.4 MB of binaries.
It is input data for the parity test of the sparse R1CS x witness kernel on real sparsity (tests/test_gpu_protocol.py) and
for bench.py's timing-only real-library prove leg.  Run in the build container only (reads /root/reference); the .xz it
writes is the committed fixture.  The reader it is built with (formats.read_r1cs) is the one
tests/test_protocol_cpu.py::test_r1cs_reader_on_the_reference_library checks against subcircuitInfo.json."""
import json
import lzma
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
from tokamak_b200.protocol import formats as F  # noqa: E402

SRC = "/root/reference/packages/frontend/qap-compiler/subcircuits/library"


def main():
    params = json.load(open(os.path.join(SRC, "setupParams.json")))
    infos = json.load(open(os.path.join(SRC, "subcircuitInfo.json")))
    coeffs, cmap, subs = [], {}, []
    for info in infos:
        r = F.read_r1cs(os.path.join(SRC, "r1cs", f"subcircuit{info['id']}.r1cs"))
        assert r.n_wires == info["Nwires"] and r.n_constraints == info["Nconsts"]
        lens, wires, cidx = [], [], []
        for abc in r.constraints:
            for lc in abc:
                lens.append(len(lc))
                for w, c in lc:
                    if c not in cmap:
                        cmap[c] = len(coeffs)
                        coeffs.append(hex(c))
                    wires.append(w)
                    cidx.append(cmap[c])
        subs.append({"id": info["id"], "n_wires": r.n_wires, "n_constraints": r.n_constraints, "lens": lens, "wires": wires, "coeff_idx": cidx})
    keep = ("id", "name", "Nwires", "Nconsts", "Out_idx", "In_idx", "flattenMap")
    doc = {"source": "packages/frontend/qap-compiler/subcircuits/library (setupParams.json, subcircuitInfo.json, r1cs/subcircuit0..13.r1cs)",
           "setupParams": params, "subcircuitInfo": [{k: i[k] for k in keep} for i in infos], "coeffs": coeffs, "r1cs": subs}
    out = os.path.join(ROOT, "tests", "golden", "real_library.json.xz")
    with lzma.open(out, "wt", preset=9) as f:
        json.dump(doc, f, separators=(",", ":"))
    print(out, os.path.getsize(out), "bytes;", sum(len(s["wires"]) for s in subs), "non-zeros;", len(coeffs), "distinct coefficients")


if __name__ == "__main__":
    main()
