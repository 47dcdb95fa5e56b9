"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch)
as plumbing.  The reference has no multi-device code at all (libs/src/utils/mod.rs:90-96 pins device 0);
this is new work scoped by SURVEY.md §8(e).

* MSM shards by point range: every rank runs the full Pippenger on its slice, the 96-byte affine partial
  sums are all-gathered and combined on rank 0 (G-1 additions).  No data-path collective besides that.
* The bivariate NTT shards by rows (X index).  The Y pass is local; the X pass needs whole columns, so the
  exchange step is one all-to-all: rank g sends the (x/G) x (y/G) tile of column block p to peer p and ends
  up with all x rows of column block g.  The result stays column-sharded (pointwise work is
  layout-agnostic); the inverse transform starts from that layout and returns to row shards.

* Fused exchange (`PeerExchange`, `bintt_sharded_*_fused`): the same re-sharding without NCCL on the data path.  The last
  pass of the local transform stores every output element straight into the destination rank's buffer over NVLink
  (peer-mapped symmetric memory; `tkm_ntt_batch_scatter`), so the transpose copy and the all-to-all disappear; ranks only
  meet at two stream-ordered barriers (before the stores: the peers have finished reading their buffers; after: all stores
  have landed).

The local compute steps are injected (`LocalOps`) so the exchange logic can be exercised on CPU with the
gloo backend in tests; the default implementation calls the CUDA library through the C-ABI.
"""
import numpy as np
import torch
import torch.distributed as dist

FORWARD, INVERSE = 0, 1


class CudaLocalOps:
    """Local steps on this rank's GPU through libtokamak_b200 (device pointers from torch tensors)."""

    def __init__(self, ctx):
        self.ctx = ctx
        # run the library on torch's current stream so kernels and NCCL collectives are ordered
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    def ntt_rows(self, t, n, batch, direction, coset=None):
        self.ctx.ntt_batch_dev(t.data_ptr(), t.data_ptr(), n, batch, False, direction, coset)

    def ntt_cols(self, t, n, batch, direction, coset=None):
        self.ctx.ntt_batch_dev(t.data_ptr(), t.data_ptr(), n, batch, True, direction, coset)

    def ntt_rows_scatter(self, t, n, batch, direction, coset, peer_ptrs, stride_a, stride_b, b0):
        self.ctx.ntt_batch_scatter(t.data_ptr(), n, batch, False, direction, coset, peer_ptrs, stride_a, stride_b, b0)

    def ntt_cols_scatter(self, t, n, batch, direction, coset, peer_ptrs, stride_a, stride_b, b0):
        self.ctx.ntt_batch_scatter(t.data_ptr(), n, batch, True, direction, coset, peer_ptrs, stride_a, stride_b, b0)

    def msm(self, scalars_t, bases_t, n):
        return self.ctx.msm_g1_dev(scalars_t.data_ptr(), False, bases_t.data_ptr(), n)

    def g1_add(self, a, b):
        return self.ctx.g1_add(a, b)

    def g1_sum(self, points):
        return self.ctx.g1_sum(points)


def shard_range(total, world, rank):
    """Contiguous point / row range of `rank`: sizes differ by at most one (ragged totals allowed)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def msm_sharded(ops, scalars_t, bases_t, n_local, group=None):
    """Sum over all ranks of MSM(local scalars, local bases).  Returns the total on rank 0 (12 x u64 affine,
    canonical) and this rank's partial elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    part = ops.msm(scalars_t, bases_t, n_local)
    if world == 1:
        return part
    dev = scalars_t.device
    t = torch.from_numpy(np.ascontiguousarray(part, dtype=np.uint64).view(np.int64).copy()).to(dev)
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t, group=group)
    if rank != 0:
        return part
    parts = torch.stack(gathered).cpu().numpy().view(np.uint64)
    if hasattr(ops, "g1_sum"):
        return ops.g1_sum(parts)  # one launch for the G-1 additions and the affine conversion
    acc = parts[0]
    for g in parts[1:]:
        acc = ops.g1_add(acc, g)
    return acc


def _exchange_rows_to_cols(t, x_local, y, world, group):
    """[x_local][y] row shard -> [world*x_local][y/world] column shard (elements are 4 x int64)."""
    yb = y // world
    send = t.view(x_local, world, yb, 4).permute(1, 0, 2, 3).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv.view(world * x_local, yb, 4)


def _exchange_cols_to_rows(t, x, y_local, world, group):
    """[x][y_local] column shard -> [x/world][world*y_local] row shard."""
    xb = x // world
    send = t.view(world, xb, y_local, 4).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    return recv.permute(1, 0, 2, 3).contiguous().view(xb, world * y_local, 4)


def bintt_sharded_forward(ops, t, x, y, coset_x=None, coset_y=None, group=None):
    """t: this rank's rows [x/G][y] (int64 view of Fr, Montgomery form on GPU).  Returns the evaluations as a
    column shard [x][y/G]: entry (k, l_local) is the value at (omega_x^k, omega_y^(g*y/G + l_local))."""
    world = dist.get_world_size(group)
    x_local = x // world
    assert x % world == 0 and y % world == 0, "x and y must be divisible by the number of ranks"
    ops.ntt_rows(t, y, x_local, FORWARD, coset_y)
    if world > 1:
        t = _exchange_rows_to_cols(t, x_local, y, world, group)
    ops.ntt_cols(t, x, y // world, FORWARD, coset_x)
    return t


def bintt_sharded_inverse(ops, t, x, y, coset_x=None, coset_y=None, group=None):
    """Inverse of bintt_sharded_forward: column shard [x][y/G] of evaluations -> row shard [x/G][y] of coefficients."""
    world = dist.get_world_size(group)
    y_local = y // world
    assert x % world == 0 and y % world == 0
    ops.ntt_cols(t, x, y_local, INVERSE, coset_x)
    if world > 1:
        t = _exchange_cols_to_rows(t, x, y_local, world, group)
    ops.ntt_rows(t, y, x // world, INVERSE, coset_y)
    return t


class PeerExchange:
    """Peer-mapped staging buffers for the fused re-sharding of an x-by-y bivariate transform: every rank owns one
    column-shard buffer [x][y/G] and one row-shard buffer [x/G][y]; all ranks hold device pointers to all of them
    (torch symmetric memory: CUDA VMM allocations exchanged once at rendezvous)."""

    def __init__(self, x, y, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world & (self.world - 1) or x % self.world or y % self.world:
            raise ValueError("fused exchange needs a power-of-two number of ranks dividing both extents")
        self.x, self.y = x, y
        n_local = x * y // self.world
        self.cols = symm_mem.empty((n_local, 4), dtype=torch.int64, device=device)
        self.rows = symm_mem.empty((n_local, 4), dtype=torch.int64, device=device)
        self.h_cols = symm_mem.rendezvous(self.cols, group)
        self.h_rows = symm_mem.rendezvous(self.rows, group)
        self.cols_ptrs = [int(p) for p in self.h_cols.buffer_ptrs]
        self.rows_ptrs = [int(p) for p in self.h_rows.buffer_ptrs]


def bintt_sharded_forward_fused(ops, ex, t, coset_x=None, coset_y=None):
    """bintt_sharded_forward with the exchange fused into the Y pass: rows [x/G][y] in `t` -> evaluations as the column
    shard [x][y/G] in ex.cols (returned as a view)."""
    x, y, world, rank = ex.x, ex.y, ex.world, ex.rank
    x_local, yb = x // world, y // world
    ex.h_cols.barrier()  # nobody is still reading a column buffer from an earlier transform
    # output (row r, column l) of this rank -> peer l // yb, element (rank*x_local + r) * yb + l % yb
    ops.ntt_rows_scatter(t, y, x_local, FORWARD, coset_y, ex.cols_ptrs, 1, yb, rank * x_local)
    ex.h_cols.barrier()  # every rank's stores have landed
    ops.ntt_cols(ex.cols, x, yb, FORWARD, coset_x)
    return ex.cols.view(x, yb, 4)


def bintt_sharded_inverse_fused(ops, ex, t, coset_x=None, coset_y=None):
    """Inverse of the above: column shard [x][y/G] in `t` (may be ex.cols) -> coefficients as the row shard [x/G][y] in
    ex.rows."""
    x, y, world, rank = ex.x, ex.y, ex.world, ex.rank
    xb, yb = x // world, y // world
    ex.h_rows.barrier()
    # output (row k, column c) of this rank -> peer k // xb, element (k % xb) * y + rank*yb + c
    ops.ntt_cols_scatter(t, x, yb, INVERSE, coset_x, ex.rows_ptrs, y, 1, rank * yb)
    ex.h_rows.barrier()
    ops.ntt_rows(ex.rows, y, xb, INVERSE, coset_y)
    return ex.rows.view(xb, y, 4)
