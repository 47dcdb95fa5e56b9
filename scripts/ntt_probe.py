"""Profiling target: a few device-resident biNTT 16384x512 transforms (forward and inverse)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
import tokamak_b200 as T  # noqa: E402

x, y = 16384, 512
ctx = T.Context(0)
ctx.init_ntt_domain_for_size(x * y)
rng = np.random.default_rng(1)
a = rng.integers(0, 1 << 62, size=(x * y, 4), dtype=np.uint64)
d = ctx.upload_fr(a, to_mont=False)
for _ in range(3):
    ctx.bintt_dev(d, d, x, y, T.FORWARD)
    ctx.bintt_dev(d, d, x, y, T.INVERSE)
ctx.time_begin()
for _ in range(4):
    ctx.bintt_dev(d, d, x, y, T.FORWARD)
print("forward ms", ctx.time_end() / 4)
for k, nm in enumerate(["IMAD.U32", "IMAD.WIDE.U32", "Fr mul", "Fq mul", "XYZZ madd"]):
    print(f"microbench {nm}: {ctx.microbench(k)/1e9:.2f} Gops/s")
ctx.close()
