"""Developer probe: MSM time against the bucket-segment length (TKM_MSM_LOGG), the accumulation chunk target
(TKM_MSM_CHUNK) and the slices per (window, bit) of the window reduction (TKM_MSM_SPLITS) at the prover's commitment sizes.  The knobs are read at every call."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tokamak-zk-evm_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import oracle_ffi as O  # noqa: E402  (input generation only)
import pyref as P  # noqa: E402
import tokamak_b200 as T  # noqa: E402

KNOBS = ("TKM_MSM_LOGG", "TKM_MSM_CHUNK", "TKM_MSM_SPLITS")
ctx = T.Context(0)
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
out = {}
for logn in (18, 20, 22):
    n = 1 << logn
    ks, ss = O.random_fr(1000 + logn, n), O.random_fr(2000 + logn, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dp = ctx.dev_alloc(n * 96)
    T.check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, ctypes.c_void_p(dk), 0, n, ctypes.c_void_p(dp)))
    T.check(ctx.lib.tkm_g1_bases_to_mont(ctx.h, ctypes.c_void_p(dp), ctypes.c_void_p(dp), n))
    ds = ctx.upload_fr(ss, to_mont=False)
    for k in KNOBS:
        os.environ.pop(k, None)
    ref = ctx.msm_g1_dev(ds, False, dp, n)
    row = {}
    for knob, vals in (("TKM_MSM_LOGG", (None, 3, 4, 5)), ("TKM_MSM_CHUNK", (None, 192, 256, 384)), ("TKM_MSM_SPLITS", (None, 1, 2, 4, 8))):
        for v in vals:
            for k in KNOBS:
                os.environ.pop(k, None)
            if v is not None:
                os.environ[knob] = str(v)
            got = ctx.msm_g1_dev(ds, False, dp, n)
            assert np.array_equal(got, ref), (logn, knob, v)
            ctx.msm_g1_dev(ds, False, dp, n)
            ctx.time_begin()
            for _ in range(4):
                ctx.msm_g1_dev(ds, False, dp, n)
            row[f"{knob[8:]}={v}"] = round(ctx.time_end() / 4, 3)
    out[logn] = row
    print(logn, row, flush=True)
    for p_ in (dk, dp, ds):
        ctx.dev_free(p_)
print(json.dumps(out))
