"""BASELINE.json config 2: standalone G1 MSM sweep 2^16..2^24, random scalars, bases k_i*G generated on the device,
result checked bit-exactly against (sum s_i k_i)*G computed by the CPU oracle, then timed (device-resident)."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tokamak-zk-evm_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import oracle_ffi as O  # noqa: E402  (checker only)
import pyref as P  # noqa: E402
import tokamak_b200 as T  # noqa: E402

ctx = T.Context(0)
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
rows = []
for logn in ([int(v) for v in os.environ["TKM_SWEEP_LOGN"].split(",")] if os.environ.get("TKM_SWEEP_LOGN") else range(16, 25)):
    n = 1 << logn
    ks, ss = O.random_fr(1000 + logn, n), O.random_fr(2000 + logn, n)
    dk = ctx.upload_fr(ks, to_mont=False)
    dp = ctx.dev_alloc(n * 96)
    T.check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, ctypes.c_void_p(dk), 0, n, ctypes.c_void_p(dp)))
    T.check(ctx.lib.tkm_g1_bases_to_mont(ctx.h, ctypes.c_void_p(dp), ctypes.c_void_p(dp), n))
    ds = ctx.upload_fr(ss, to_mont=False)
    got = ctx.msm_g1_dev(ds, False, dp, n)
    exp = O.g1_mul(G, O.fr_inner_product(ss, ks))
    ok = bool(np.array_equal(got, exp))
    for _ in range(2):
        ctx.msm_g1_dev(ds, False, dp, n)
    reps = 5 if logn <= 22 else 3
    ctx.time_begin()
    for _ in range(reps):
        ctx.msm_g1_dev(ds, False, dp, n)
    ms = ctx.time_end() / reps
    rows.append({"log2_n": logn, "ms": round(ms, 3), "mpts_per_s": round(n / ms / 1e3, 2), "bit_exact_vs_oracle": ok})
    print(rows[-1], flush=True)
    for p_ in (dk, dp, ds):
        ctx.dev_free(p_)
print(json.dumps({"msm_sweep": rows}))
