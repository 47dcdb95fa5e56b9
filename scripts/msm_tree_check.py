"""Developer check of the affine pair tree: MSM results with and without it, per-size timing and level counts."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tokamak_b200 as T
import oracle_ffi as O
import pyref as P

O.build()
ctx = T.Context(0)
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
for logn in [int(a) for a in sys.argv[1:]] or [12, 16, 20, 22]:
    n = 1 << logn
    ks, ss = O.random_fr(80 + logn, n), O.random_fr(81 + logn, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dp = ctx.dev_alloc(n * 96)
    ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dp)
    ctx.lib.tkm_g1_bases_to_mont(ctx.h, dp, dp, n)
    ds = ctx.upload_fr(ss, to_mont=False)
    exp = O.g1_mul(G, O.fr_inner_product(ss, ks))
    for lv in (None, "0") + tuple(os.environ.get("TREE_TRY", "").split()):
        if lv is None:
            os.environ.pop("TKM_MSM_TREE_LEVELS", None)
        else:
            os.environ["TKM_MSM_TREE_LEVELS"] = lv
        got = ctx.msm_g1_dev(ds, False, dp, n)
        ok = np.array_equal(got, exp)
        for _ in range(2):
            ctx.msm_g1_dev(ds, False, dp, n)
        ctx.sync(); ctx.time_begin()
        for _ in range(5):
            ctx.msm_g1_dev(ds, False, dp, n)
        ms = ctx.time_end() / 5
        print(f"2^{logn} tree={lv or 'auto'} ok={ok} {ms:.3f} ms {n/ms/1e3:.1f} Mpts/s acc_phase={ctx.kernel_time_last():.3f} ms stats={ctx.msm_tree_stats()}", flush=True)
    for p_ in (dk, dp, ds):
        ctx.dev_free(p_)
ctx.close()
