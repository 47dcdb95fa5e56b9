import sys, os, cProfile, pstats, io, time
ROOT="/root/repo"
for p in (os.path.join(ROOT,"tokamak-zk-evm_b200"), os.path.join(ROOT,"oracle"), os.path.join(ROOT,"scripts")):
    sys.path.insert(0,p)
import numpy as np
import tokamak_b200 as T
from tokamak_b200.protocol import synthetic as S, setup as ST, prover as PV, qap, formats as F0
from tokamak_b200.protocol.backend import GpuBackend
ctx=T.Context(0); be=GpuBackend(ctx)
params, infos, r1cs = S.make_library(S.reference_shape())
pl, perm, inst = S.synthesize(params, infos, r1cs, small_value_fraction=0.5)
for p_ in pl:
    p_.variables = F0.ScalarArray(np.frombuffer(b"".join(v.to_bytes(32,"little") for v in p_.variables), dtype=np.uint64).reshape(-1,4))
sigma = ST.generate(be, params, infos, r1cs, ST.Tau.gen_fixed()); ctx.sync()
be.reserve(24<<30)
csr = qap.LibraryCSR(r1cs)
def one():
    pv = PV.Prover(be, params, infos, r1cs, sigma, pl, perm, inst, mixer=PV.Mixer.fixed(), library_csr=csr)
    return PV.prove(pv)
for _ in range(3): one()
t=time.perf_counter(); one(); print("plain", time.perf_counter()-t)
pr=cProfile.Profile(); pr.enable(); one(); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:5000])
