"""Synthetic qap-compiler library + synthesizer output of the reference's shapes.

The reference's real inputs cannot be produced here: witnesses come from circom WASM witness calculators driven by the
TypeScript synthesizer (no node in this image) and `placementVariables.json`/`permutation.json`/`instance.json` are not in
the tree.  This module writes a library with the same file formats, wire partition and buffer conventions
as packages/frontend/qap-compiler/subcircuits/library (public wires [0,l) fed through the four public buffers placed at
columns 0..3, interface wires [l,l_D), private wires [l_D,m_D); flattenMap = const, outputs, inputs, internals) whose
R1CS constraints are forward-solvable, so a satisfying witness with uniformly random field values can be computed
directly.  `reference_shape()` reproduces setupParams.json of the in-tree library (n = 4096, s_max = 256, m_I = 4096,
l = 728, l_free = 128, l_user = 85, l_user_out = 65, 14 subcircuits)."""
import random
from dataclasses import dataclass
from typing import List

from .formats import R1CS, Instance, Permutation, PlacementVariables, SetupParams, SubcircuitInfo
from .fr import R_MOD


@dataclass
class ComputeSpec:
    name: str
    n_out: int
    n_in: int
    n_constraints: int  # internal wires = n_constraints - n_out (every constraint defines one wire)


@dataclass
class LibrarySpec:
    n: int
    s_max: int
    m_i: int
    l_user_out: int
    l_user: int
    l_free: int
    l: int
    n_prv_in: int
    compute: List[ComputeSpec]


def reference_shape():
    """Shape parameters of the checked-in library (setupParams.json / subcircuitInfo.json: names, In/Out counts, Nconsts)."""
    return LibrarySpec(n=4096, s_max=256, m_i=4096, l_user_out=65, l_user=85, l_free=128, l=728, n_prv_in=1060, compute=[
        ComputeSpec("ALU1", 2, 5, 2271), ComputeSpec("ALU2", 2, 7, 2782), ComputeSpec("DecToBit", 256, 2, 258),
        ComputeSpec("SubExpBatch", 4, 36, 3936), ComputeSpec("Accumulator", 2, 64, 329), ComputeSpec("Poseidon", 2, 15, 3793),
        ComputeSpec("JubjubExpBatch", 8, 136, 3464), ComputeSpec("EdDsaVerify", 1, 12, 26), ComputeSpec("VerifyMerkleProof", 1, 21, 3878)])


def tiny_shape():
    return LibrarySpec(n=16, s_max=8, m_i=64, l_user_out=2, l_user=4, l_free=8, l=12, n_prv_in=3, compute=[
        ComputeSpec("ALU1", 2, 3, 9), ComputeSpec("Poseidon", 1, 2, 16), ComputeSpec("Accumulator", 1, 4, 5)])


def _copy_buffer(n_wires_side):
    """out_k = in_k for k < n: wires 0 const, 1..n outs, n+1..2n ins; constraint in_k * 1 = out_k."""
    cons = [([(1 + n_wires_side + k, 1)], [(0, 1)], [(1 + k, 1)]) for k in range(n_wires_side)]
    return R1CS(1 + 2 * n_wires_side, n_wires_side, cons)


def _compute_r1cs(spec: ComputeSpec, rng):
    n_int = spec.n_constraints - spec.n_out
    assert n_int >= 0
    n_wires = 1 + spec.n_out + spec.n_in + n_int
    first_in = 1 + spec.n_out
    first_int = first_in + spec.n_in
    known = [0] + list(range(first_in, first_in + spec.n_in))
    targets = list(range(first_int, n_wires)) + list(range(1, 1 + spec.n_out))
    cons = []

    def lin(k):
        return [(w, rng.randrange(1, R_MOD)) for w in rng.sample(known, min(k, len(known)))]

    for t in targets:
        a, b = lin(rng.randint(1, 3)), lin(rng.randint(1, 2))
        c = [(t, 1)] + (lin(1) if rng.random() < 0.3 else [])
        cons.append((a, b, c))
        known.append(t)
    return R1CS(n_wires, spec.n_constraints, cons)


def make_library(spec: LibrarySpec, seed=1):
    """-> (SetupParams, [SubcircuitInfo], [R1CS])."""
    rng = random.Random(seed)
    l, m_i = spec.l, spec.m_i
    m_block, m_function = spec.l_free - spec.l_user, spec.l - spec.l_free
    infos, r1cs = [], []
    nxt_if = [l]       # next free interface wire
    nxt_prv = [l + m_i]  # next free private wire

    def take_if(k):
        s = nxt_if[0]
        nxt_if[0] += k
        if nxt_if[0] > l + m_i:
            raise ValueError("interface wires exceed m_I")
        return list(range(s, s + k))

    def add(name, r, n_out, n_in, fmap):
        infos.append(SubcircuitInfo(len(infos), name, r.n_wires, r.n_constraints, [1, n_out], [1 + n_out, n_in], fmap))
        r1cs.append(r)

    # public buffers: the public side sits in [0,l), the other side on interface wires
    k = spec.l_user_out
    add("bufferPubOut", _copy_buffer(k), k, k, take_if(1) + list(range(0, k)) + take_if(k))
    k = spec.l_user - spec.l_user_out
    c = take_if(1)
    add("bufferPubIn", _copy_buffer(k), k, k, c + take_if(k) + list(range(spec.l_user_out, spec.l_user)))
    c = take_if(1)
    add("bufferBlockIn", _copy_buffer(m_block), m_block, m_block, c + take_if(m_block) + list(range(spec.l_user, spec.l_free)))
    c = take_if(1)
    add("bufferEVMIn", _copy_buffer(m_function), m_function, m_function, c + take_if(m_function) + list(range(spec.l_free, l)))
    k = spec.n_prv_in
    c = take_if(1)
    add("bufferPrvIn", _copy_buffer(k), k, k, c + take_if(k) + take_if(k))
    for cs in spec.compute:
        r = _compute_r1cs(cs, rng)
        n_int = r.n_wires - 1 - cs.n_out - cs.n_in
        fmap = take_if(1) + take_if(cs.n_out) + take_if(cs.n_in) + list(range(nxt_prv[0], nxt_prv[0] + n_int))
        nxt_prv[0] += n_int
        add(cs.name, r, cs.n_out, cs.n_in, fmap)
    params = SetupParams(l_free=spec.l_free, l=l, l_user_out=spec.l_user_out, l_user=spec.l_user, l_D=l + m_i, m_D=nxt_prv[0], n=spec.n,
                         s_D=len(infos), s_max=spec.s_max)
    params.validate()
    return params, infos, r1cs


def solve_witness(r: R1CS, info: SubcircuitInfo, inputs):
    """Forward evaluation of a library subcircuit: wire 0 = 1, inputs given, every constraint defines the wire with
    coefficient 1 in C that is not known yet."""
    w = [None] * r.n_wires
    w[0] = 1
    i0, n_in = info.In_idx
    assert len(inputs) == n_in
    for k, v in enumerate(inputs):
        w[i0 + k] = v % R_MOD
    for a, b, c in r.constraints:
        av = sum(cf * w[x] for x, cf in a) % R_MOD
        bv = sum(cf * w[x] for x, cf in b) % R_MOD
        unknown = [(x, cf) for x, cf in c if w[x] is None]
        assert len(unknown) == 1 and unknown[0][1] == 1, "library constraint is not forward-solvable"
        rest = sum(cf * w[x] for x, cf in c if w[x] is not None) % R_MOD
        w[unknown[0][0]] = (av * bv - rest) % R_MOD
    assert all(v is not None for v in w)
    return w


def check_r1cs(r: R1CS, w):
    for a, b, c in r.constraints:
        av = sum(cf * w[x] for x, cf in a) % R_MOD
        bv = sum(cf * w[x] for x, cf in b) % R_MOD
        cv = sum(cf * w[x] for x, cf in c) % R_MOD
        if av * bv % R_MOD != cv:
            return False
    return True


def fill_witness_unchecked(seed=7):
    """A stand-in for the witness calculators the real library needs (circom wasm, not runnable here): wire 0 = 1, the
    inputs as given, every other wire a seeded random field element.  The constraints are NOT satisfied -- for timing-only
    proves and for kernels whose result does not depend on satisfiability (the sparse R1CS x witness products)."""
    rng = random.Random(seed)

    def solve(r: R1CS, info: SubcircuitInfo, inputs):
        w = [rng.randrange(R_MOD) for _ in range(r.n_wires)]
        w[0] = 1
        i0, n_in = info.In_idx
        assert len(inputs) == n_in
        for k, v in enumerate(inputs):
            w[i0 + k] = v % R_MOD
        return w

    return solve


def synthesize(params: SetupParams, infos, r1cs, n_placements=None, seed=2, small_value_fraction=0.0, solver=None):
    """A random dataflow over the library: -> (placements, permutation, instance).  `solver` replaces the forward
    evaluation of a subcircuit (solve_witness) -- fill_witness_unchecked for libraries that are not forward-solvable.

    Columns 0..4 are the five buffers; every later column is a compute subcircuit whose inputs are copies of values
    produced earlier (outputs of the input buffers or of earlier subcircuits); bufferPubOut's inputs copy subcircuit
    outputs.  Copy constraints are emitted as cycles over (interface wire, placement) nodes (permutation.json)."""
    rng = random.Random(seed)
    solve_witness = solver or globals()["solve_witness"]
    s_max, l = params.s_max, params.l
    n_pl = s_max if n_placements is None else n_placements
    assert 6 <= n_pl <= s_max

    def rnd():
        if rng.random() < small_value_fraction:
            return rng.randrange(0, 1 << 16)
        return rng.randrange(R_MOD)

    byname = {s.name: s for s in infos}
    sources = []   # (value, global wire, placement)
    classes = {}   # source index -> list of nodes (global wire, placement) that copy it
    placements = [None] * n_pl

    def place_input_buffer(col, name):
        info = byname[name]
        vals = [rnd() for _ in range(info.In_idx[1])]
        w = solve_witness(r1cs[info.id], info, vals)
        placements[col] = PlacementVariables(info.id, w)
        o0, n_out = info.Out_idx
        for k in range(n_out):
            sources.append((w[o0 + k], info.flattenMap[o0 + k], col))
        return vals

    user_in = place_input_buffer(1, "bufferPubIn")
    block_in = place_input_buffer(2, "bufferBlockIn")
    function_in = place_input_buffer(3, "bufferEVMIn")
    place_input_buffer(4, "bufferPrvIn")
    compute = [s for s in infos if not s.name.startswith("buffer")]
    compute_outputs = []
    for col in range(5, n_pl):
        info = compute[(col - 5) % len(compute)] if col - 5 < len(compute) else rng.choice(compute)
        i0, n_in = info.In_idx
        picks = [rng.randrange(len(sources)) for _ in range(n_in)]
        w = solve_witness(r1cs[info.id], info, [sources[p][0] for p in picks])
        placements[col] = PlacementVariables(info.id, w)
        for k, p in enumerate(picks):
            classes.setdefault(p, []).append((info.flattenMap[i0 + k], col))
        o0, n_out = info.Out_idx
        for k in range(n_out):
            compute_outputs.append(len(sources))
            sources.append((w[o0 + k], info.flattenMap[o0 + k], col))
    info = byname["bufferPubOut"]
    i0, n_in = info.In_idx
    picks = [rng.choice(compute_outputs) for _ in range(n_in)]
    w = solve_witness(r1cs[info.id], info, [sources[p][0] for p in picks])
    placements[0] = PlacementVariables(info.id, w)
    for k, p in enumerate(picks):
        classes.setdefault(p, []).append((info.flattenMap[i0 + k], 0))
    user_out = [w[info.Out_idx[0] + k] for k in range(info.Out_idx[1])]

    permutation = []
    for p, nodes in classes.items():
        cyc = [(sources[p][1], sources[p][2])] + nodes
        for a, b in zip(cyc, cyc[1:] + cyc[:1]):
            permutation.append(Permutation(row=a[0] - l, col=a[1], X=b[0] - l, Y=b[1]))
    instance = Instance(a_pub_user=user_out + user_in, a_pub_block=block_in, a_pub_function=function_in)
    return placements, permutation, instance
