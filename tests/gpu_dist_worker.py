"""torchrun worker (one process per GPU, NCCL): sharded biNTT and sharded MSM against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tokamak-zk-evm_b200")):
    sys.path.insert(0, p)
import oracle_ffi as O  # noqa: E402
import pyref as P  # noqa: E402
import tokamak_b200 as T  # noqa: E402
from tokamak_b200 import dist as D  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
ctx = T.Context(local_rank)
ctx.init_ntt_domain_for_size(1 << 20)
ops = D.CudaLocalOps(ctx)
dev = torch.device("cuda", local_rank)


def to_dev_mont(a):
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(dev)
    T.check(ctx.lib.tkm_fr_to_mont(ctx.h, t.data_ptr(), t.data_ptr(), t.numel() // 4))
    return t


def from_dev_mont(t):
    t = t.contiguous().clone()
    T.check(ctx.lib.tkm_fr_from_mont(ctx.h, t.data_ptr(), t.data_ptr(), t.numel() // 4))
    torch.cuda.synchronize()
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


for (x, y, cx, cy) in ((64, 32, None, None), (2048, 64, 12345, 678910), (4096, 256, None, None)):
    full = O.random_fr(11 + x, x * y)
    lo, hi = D.shard_range(x, world, rank)
    t = to_dev_mont(full.reshape(x, y, 4)[lo:hi].copy()).view(hi - lo, y, 4)
    ev = D.bintt_sharded_forward(ops, t, x, y, cx, cy)
    exp = O.bintt(full, x, y, False, None if cx is None else O.fr_from_int(cx), None if cy is None else O.fr_from_int(cy)).reshape(x, y, 4)
    yb = y // world
    got = from_dev_mont(ev).reshape(x, yb, 4)
    assert np.array_equal(got, exp[:, rank * yb:(rank + 1) * yb]), f"forward column shard {x}x{y}"
    back = D.bintt_sharded_inverse(ops, ev.contiguous(), x, y, cx, cy)
    assert np.array_equal(from_dev_mont(back).reshape(hi - lo, y, 4), full.reshape(x, y, 4)[lo:hi]), "round trip"

# fused exchange: the last NTT pass stores into the peers' buffers over NVLink (no transpose copy, no all-to-all)
if world & (world - 1) == 0:
    for (x, y, cx, cy) in ((64, 32, None, None), (2048, 64, 12345, 678910), (4096, 256, None, None), (16384, 512, 7, None)):
        if x % world or y % world:
            continue
        ctx.init_ntt_domain_for_size(max(1 << 20, x * y))
        ex = D.PeerExchange(x, y, dev)
        full = O.random_fr(31 + x, x * y)
        lo, hi = D.shard_range(x, world, rank)
        t = to_dev_mont(full.reshape(x, y, 4)[lo:hi].copy()).view(hi - lo, y, 4)
        exp = O.bintt(full, x, y, False, None if cx is None else O.fr_from_int(cx), None if cy is None else O.fr_from_int(cy)).reshape(x, y, 4)
        yb = y // world
        for rep in range(2):  # twice: the barriers must also protect buffer reuse
            ev = D.bintt_sharded_forward_fused(ops, ex, t, cx, cy)
            assert np.array_equal(from_dev_mont(ev).reshape(x, yb, 4), exp[:, rank * yb:(rank + 1) * yb]), f"fused forward {x}x{y}"
            back = D.bintt_sharded_inverse_fused(ops, ex, ev, cx, cy)
            assert np.array_equal(from_dev_mont(back).reshape(hi - lo, y, 4), full.reshape(x, y, 4)[lo:hi]), f"fused round trip {x}x{y}"
        del ex

n = 50001  # ragged shards
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
pts = O.g1_fixed_base_mul_batch(G, O.random_fr(21, n))
ss = O.random_fr(22, n)
lo, hi = D.shard_range(n, world, rank)
d_b = torch.from_numpy(pts[lo:hi].copy().view(np.int64)).to(dev)
T.check(ctx.lib.tkm_g1_bases_to_mont(ctx.h, d_b.data_ptr(), d_b.data_ptr(), hi - lo))
d_s = torch.from_numpy(ss[lo:hi].copy().view(np.int64)).to(dev)
tot = D.msm_sharded(ops, d_s, d_b, hi - lo)
if rank == 0:
    assert np.array_equal(tot, O.msm_g1(ss, pts)), "sharded MSM total"
# ---- the same sharded paths through the C-ABI's own communicator (tkm_comm_*: NCCL unique-id bootstrap, no torch on the
# data path): what a Rust prover binds.  The id travels out of band (here: a torch broadcast).
import ctypes  # noqa: E402

torch.cuda.synchronize()
ctx.set_stream(None)  # back to the context's own stream
idbuf = np.zeros(128, dtype=np.uint8)
if rank == 0:
    T.check(ctx.lib.tkm_comm_unique_id(idbuf.ctypes.data_as(ctypes.c_void_p)))
idt = torch.from_numpy(idbuf).to(dev)
dist.broadcast(idt, 0)
idbuf = idt.cpu().numpy().copy()
T.check(ctx.lib.tkm_comm_init(ctx.h, idbuf.ctypes.data_as(ctypes.c_void_p), rank, world))
r_, w_ = ctypes.c_int32(), ctypes.c_int32()
T.check(ctx.lib.tkm_comm_rank(ctx.h, ctypes.byref(r_), ctypes.byref(w_)))
assert (r_.value, w_.value) == (rank, world)
out = np.zeros(12, dtype=np.uint64)
T.check(ctx.lib.tkm_msm_g1_sharded(ctx.h, d_s.data_ptr(), 0, d_b.data_ptr(), hi - lo, out.ctypes.data_as(ctypes.c_void_p)))
assert np.array_equal(out, O.msm_g1(ss, pts)), "tkm_msm_g1_sharded: every rank must hold the total"
for (x, y, cx, cy) in ((64, 32, None, None), (2048, 64, 12345, 678910), (16384, 512, None, None)):
    ctx.init_ntt_domain_for_size(max(1 << 20, x * y))
    full = O.random_fr(51 + x, x * y)
    xb, yb = x // world, y // world
    rows = torch.from_numpy(np.ascontiguousarray(full.reshape(x, y, 4)[rank * xb:(rank + 1) * xb]).view(np.int64)).to(dev)
    cols = torch.empty_like(rows)
    back = torch.empty_like(rows)
    torch.cuda.synchronize()  # the library runs on its own stream from here on
    T.check(ctx.lib.tkm_fr_to_mont(ctx.h, rows.data_ptr(), rows.data_ptr(), rows.numel() // 4))
    _kx, bx = T.fr_bytes(cx)
    _ky, by = T.fr_bytes(cy)
    T.check(ctx.lib.tkm_bintt_sharded(ctx.h, rows.data_ptr(), cols.data_ptr(), x, y, 0, bx, by))
    T.check(ctx.lib.tkm_bintt_sharded(ctx.h, cols.data_ptr(), back.data_ptr(), x, y, 1, bx, by))
    T.check(ctx.lib.tkm_fr_from_mont(ctx.h, cols.data_ptr(), cols.data_ptr(), cols.numel() // 4))
    T.check(ctx.lib.tkm_fr_from_mont(ctx.h, back.data_ptr(), back.data_ptr(), back.numel() // 4))
    ctx.sync()
    host = lambda t_: t_.cpu().numpy().view(np.uint64).reshape(-1, 4)
    exp = O.bintt(full, x, y, False, None if cx is None else O.fr_from_int(cx), None if cy is None else O.fr_from_int(cy)).reshape(x, y, 4)
    assert np.array_equal(host(cols).reshape(x, yb, 4), exp[:, rank * yb:(rank + 1) * yb]), f"tkm_bintt_sharded forward {x}x{y}"
    assert np.array_equal(host(back).reshape(xb, y, 4), full.reshape(x, y, 4)[rank * xb:(rank + 1) * xb]), f"tkm_bintt_sharded round trip {x}x{y}"
T.check(ctx.lib.tkm_comm_destroy(ctx.h))
dist.barrier()
sys.stdout.write(f"rank{rank}of{world}ok\n")  # one write call: no interleaving between ranks
sys.stdout.flush()
ctx.close()
dist.destroy_process_group()
