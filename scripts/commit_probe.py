"""Profiling target: commitments of a 4097x257 polynomial against the 8192x512 CRS, plain and with fixed-base tables."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import prove_replay as R  # noqa: E402
import tokamak_b200 as T  # noqa: E402

ctx = T.Context(0)
sigma, table = R.make_sigma(ctx)
rep = R.Replay(ctx)
p = rep.poly(4097, 257)
c = int(os.environ.get("PRE_C", "0"))
if c:
    sigma.precompute(c)
for _ in range(3):
    sigma.encode_poly(p)
ctx.sync()
t0 = time.perf_counter()
for _ in range(5):
    sigma.encode_poly(p)
ctx.sync()
print(f"PRE_C={c}: commit 4097x257 = {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")
