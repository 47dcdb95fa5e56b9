"""Developer timing probe (not the graded bench): microbenchmarks + device-resident biNTT / MSM timings."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
import tokamak_b200 as T  # noqa: E402


def main():
    ctx = T.Context(0)
    names = ["IMAD.U32", "IMAD.WIDE.U32", "Fr mul", "Fq mul", "XYZZ madd"]
    for k, nm in enumerate(names):
        v = ctx.microbench(k)
        print(f"microbench {nm}: {v/1e9:.2f} Gops/s", flush=True)
    ctx.init_ntt_domain_for_size(1 << 23)
    rng = np.random.default_rng(1)
    for (x, y) in ((4096, 256), (8192, 512), (16384, 512)):
        n = x * y
        a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        d = ctx.upload_fr(a, to_mont=False)
        for direction in (T.FORWARD, T.INVERSE):
            for _ in range(3):
                ctx.bintt_dev(d, d, x, y, direction)
            ctx.time_begin()
            reps = 5
            for _ in range(reps):
                ctx.bintt_dev(d, d, x, y, direction)
            ms = ctx.time_end() / reps
            print(f"biNTT {x}x{y} dir={direction}: {ms:.3f} ms  {n/ms/1e6:.3f} Gelem/s  hbm-frac {128*n/(ms*1e-3)/6541.8e9:.4f}", flush=True)
        ctx.dev_free(d)
    # MSM: distinct addresses, 4096 distinct random points tiled
    G = np.zeros(12, dtype=np.uint64)
    gx = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
    gy = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
    G = np.frombuffer(gx.to_bytes(48, "little") + gy.to_bytes(48, "little"), dtype=np.uint64).copy()
    for logn in (16, 18, 20, 22):
        n = 1 << logn
        ks = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        ss = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        ss[:, 3] >>= 0
        dk = ctx.upload_fr(ks, to_mont=False)
        dp = ctx.dev_alloc(n * 96)
        t0 = time.time()
        T.check(ctx.lib.tkm_g1_fixed_base_mul(ctx.h, G.ctypes.data, dk, 0, n, dp))
        ctx.sync()
        print(f"fixed-base mul 2^{logn}: {time.time()-t0:.3f} s", flush=True)
        T.check(ctx.lib.tkm_g1_bases_to_mont(ctx.h, dp, dp, n))
        ds = ctx.upload_fr(ss, to_mont=False)
        ctx.msm_g1_dev(ds, False, dp, n)
        ctx.time_begin()
        reps = 3
        for _ in range(reps):
            ctx.msm_g1_dev(ds, False, dp, n)
        ms = ctx.time_end() / reps
        print(f"MSM 2^{logn}: {ms:.3f} ms  {n/ms/1e3:.2f} Mpts/s", flush=True)
        for p in (dk, dp, ds):
            ctx.dev_free(p)
    ctx.close()


if __name__ == "__main__":
    main()
