"""Developer probe: Fr butterfly stream on the fused-chain reduction (kind 6) against the round-1 add-chain reduction (kind 7),
the Fr product stream, and the device-resident biNTT at the prover's shapes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
import tokamak_b200 as T  # noqa: E402

ctx = T.Context(0)
for kind, nm in ((1, "IMAD.WIDE.U32"), (5, "IMAD.WIDE.U32.X chains"), (2, "Fr mul"), (6, "Fr butterfly (fused-chain reduction)"), (7, "Fr butterfly (round-1 add chains)")):
    best = max(ctx.microbench(kind) for _ in range(3))
    print(f"microbench {nm}: {best / 1e9:.2f} G/s", flush=True)
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
    print(f"biNTT {x}x{y}: forward {res['forward']:.4f} ms, inverse {res['inverse']:.4f} ms, {x * y / res['forward'] / 1e6:.2f} Gelem/s", flush=True)
    ctx.dev_free(d)
ctx.close()
