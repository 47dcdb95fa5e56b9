"""Developer probe: host-buffer MSM (tkm_msm_g1_host, pinned buffers) against the number of copy/compute pieces
(TKM_MSM_HOST_PIECES) at 2^22 points."""
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tokamak-zk-evm_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import oracle_ffi as O  # noqa: E402  (input generation only)
import pyref as P  # noqa: E402
import tokamak_b200 as T  # noqa: E402
import torch  # noqa: E402

ctx = T.Context(0)
n = 1 << 22
G = np.frombuffer(P.g1_to_bytes(P.G1_GEN), dtype=np.uint64).copy()
ks, ss = O.random_fr(1022, n), O.random_fr(2022, n)
dk = ctx.upload_fr(ks, to_mont=False)
dp = ctx.dev_alloc(n * 96)
T.check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, ctypes.c_void_p(dk), 0, n, ctypes.c_void_p(dp)))
pts = torch.empty((n, 12), dtype=torch.int64, pin_memory=True)
sc = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
T.check(ctx.lib.tkm_memcpy_d2h(ctx.h, pts.data_ptr(), dp, n * 96))
sc.numpy().view(np.uint64)[:] = ss
import ctypes as C
out = np.zeros(12, dtype=np.uint64)
ref = None
for pieces in (None, 1, 2, 3, 4, 5, 6, 8):
    if pieces is None:
        os.environ.pop("TKM_MSM_HOST_PIECES", None)
    else:
        os.environ["TKM_MSM_HOST_PIECES"] = str(pieces)
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        T.check(ctx.lib.tkm_msm_g1_host(ctx.h, sc.data_ptr(), pts.data_ptr(), n, out.ctypes.data))
        ts.append(time.perf_counter() - t0)
    if ref is None:
        ref = out.copy()
    assert np.array_equal(out, ref)
    print(pieces, [round(t * 1e3, 2) for t in ts], flush=True)
