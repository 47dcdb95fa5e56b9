"""Developer probe: commitment of an 8192 x 512 polynomial on a device-resident CRS grid, plain and with fixed-base tables of
several window widths (tkm_crs_precompute); result compared with the known-discrete-log identity."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tokamak-zk-evm_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import oracle_ffi as O  # noqa: E402  (checker only)
import pyref as P  # noqa: E402
import tokamak_b200 as T  # noqa: E402

ctx = T.Context(0)
rs_x, rs_y = 8192, 512
n = rs_x * rs_y
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
ks, ss = O.random_fr(610, n), O.random_fr(611, n)
dk = ctx.upload_fr(ks, to_mont=False)
dg = ctx.dev_alloc(n * 96)
T.check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, ctypes.c_void_p(dk), 0, n, ctypes.c_void_p(dg)))
h = ctypes.c_void_p()
T.check(ctx.lib.tkm_crs_from_device(ctx.h, ctypes.c_void_p(dg), rs_x, rs_y, 1, ctypes.byref(h)))
exp = O.g1_mul(G, O.fr_inner_product(ss, ks))
poly = T.DensePolynomialExt.from_coeffs(ctx, ss, rs_x, rs_y)
out = np.zeros(12, dtype=np.uint64)


def timed(label):
    T.check(ctx.lib.tkm_poly_commit(ctx.h, poly.h, h, out.ctypes.data_as(ctypes.c_void_p)))
    ok = bool(np.array_equal(out, exp))
    for _ in range(2):
        T.check(ctx.lib.tkm_poly_commit(ctx.h, poly.h, h, out.ctypes.data_as(ctypes.c_void_p)))
    ctx.time_begin()
    for _ in range(5):
        T.check(ctx.lib.tkm_poly_commit(ctx.h, poly.h, h, out.ctypes.data_as(ctypes.c_void_p)))
    ms = ctx.time_end() / 5
    lv = ctypes.c_uint32()
    cnt = (ctypes.c_uint64 * 9)()
    ctx.lib.tkm_msm_tree_stats(ctx.h, ctypes.byref(lv), cnt)
    print(f"{label}: {ms:.3f} ms  {n / ms / 1e3:.1f} Mpts/s  exact={ok}  tree levels {lv.value} entries {list(cnt)[:lv.value + 1]}", flush=True)


timed("plain (GLV, c=16)")
for c in [int(v) for v in os.environ.get("TKM_PROBE_C", "16,18,20,21,22").split(",")]:
    t0 = time.time()
    T.check(ctx.lib.tkm_crs_precompute(ctx.h, h, c))
    ctx.sync()
    tb = time.time() - t0
    timed(f"tables c={c} (built in {tb:.2f} s)")
