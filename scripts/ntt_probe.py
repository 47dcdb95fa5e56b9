"""Profiling / tuning target: device-resident biNTT transforms (forward and inverse) at the prover's shapes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
import tokamak_b200 as T  # noqa: E402

ctx = T.Context(0)
ctx.init_ntt_domain_for_size(1 << 23)
rng = np.random.default_rng(1)
for x, y in ((16384, 512), (8192, 512), (4096, 256)):
    a = rng.integers(0, 1 << 62, size=(x * y, 4), dtype=np.uint64)
    d = ctx.upload_fr(a, to_mont=False)
    for _ in range(3):
        ctx.bintt_dev(d, d, x, y, T.FORWARD)
        ctx.bintt_dev(d, d, x, y, T.INVERSE)
    res = {}
    for direction, nm in ((T.FORWARD, "forward"), (T.INVERSE, "inverse")):
        ctx.time_begin()
        for _ in range(10):
            ctx.bintt_dev(d, d, x, y, direction)
        res[nm] = ctx.time_end() / 10
    print(f"{x}x{y} tile_log={os.environ.get('TKM_NTT_TILE_LOG', 'default')}: forward {res['forward']:.4f} ms, inverse {res['inverse']:.4f} ms, "
          f"{x * y / res['forward'] / 1e6:.2f} Gelem/s")
    ctx.dev_free(d)
ctx.close()
