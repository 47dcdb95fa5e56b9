"""CPU, world_size 2, gloo: the multi-GPU sharding logic (point-range MSM shards + partial-sum combine; row-sharded
biNTT with the all-to-all transpose) with the local compute steps replaced by the CPU oracle."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = sys.argv[1]
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tokamak-zk-evm_b200")):
    sys.path.insert(0, p)
import oracle_ffi as O
import pyref as P
from tokamak_b200 import dist as D

class OracleOps:
    """Stand-in for CudaLocalOps on CPU tensors (canonical values)."""
    def _arr(self, t):
        return t.numpy().view(np.uint64).reshape(-1, 4)
    def ntt_rows(self, t, n, batch, direction, coset=None):
        a = self._arr(t)
        a[:] = O.ntt(a.copy(), n, batch, False, direction == D.INVERSE, None if coset is None else O.fr_from_int(coset))
    def ntt_cols(self, t, n, batch, direction, coset=None):
        a = self._arr(t)
        a[:] = O.ntt(a.copy(), n, batch, True, direction == D.INVERSE, None if coset is None else O.fr_from_int(coset))
    def msm(self, s, b, n):
        return O.msm_g1(s.numpy().view(np.uint64).reshape(-1, 4), b.numpy().view(np.uint64).reshape(-1, 12))
    def g1_add(self, a, b):
        return O.g1_add(a, b)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ops = OracleOps()
O.set_num_threads(2)
# ---- biNTT, row shards, plain and coset, forward then inverse
x, y = 16, 8
full = O.random_fr(5, x * y)
for cx, cy in ((None, None), (12345, 67890)):
    lo, hi = D.shard_range(x, world, rank)
    t = torch.from_numpy(full.reshape(x, y, 4)[lo:hi].copy().view(np.int64))
    ev = D.bintt_sharded_forward(ops, t, x, y, cx, cy)
    exp = O.bintt(full, x, y, False, None if cx is None else O.fr_from_int(cx), None if cy is None else O.fr_from_int(cy)).reshape(x, y, 4)
    yb = y // world
    assert np.array_equal(ev.numpy().view(np.uint64).reshape(x, yb, 4), exp[:, rank * yb:(rank + 1) * yb]), "forward column shard"
    back = D.bintt_sharded_inverse(ops, ev.contiguous(), x, y, cx, cy)
    assert np.array_equal(back.numpy().view(np.uint64).reshape(hi - lo, y, 4), full.reshape(x, y, 4)[lo:hi]), "round trip to row shard"
# ---- MSM, ragged point-range shards
n = 101
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
pts = O.g1_fixed_base_mul_batch(G, O.random_fr(6, n))
ss = O.random_fr(7, n)
lo, hi = D.shard_range(n, world, rank)
assert (lo, hi) == ((0, 51) if rank == 0 else (51, 101))
tot = D.msm_sharded(ops, torch.from_numpy(ss[lo:hi].copy().view(np.int64)), torch.from_numpy(pts[lo:hi].copy().view(np.int64)), hi - lo)
if rank == 0:
    assert np.array_equal(tot, O.msm_g1(ss, pts)), "sharded MSM total"
dist.barrier()
dist.destroy_process_group()
sys.stdout.write(f"rank{rank}ok\n")  # one write call: no interleaving between ranks
sys.stdout.flush()
'''


def test_sharded_paths_world2_gloo(tmp_path):
    import oracle_ffi as O

    O.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29653", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout


def test_shard_range_covers_everything():
    from tokamak_b200.dist import shard_range

    for total in (0, 1, 7, 8, 101, 1 << 22):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
