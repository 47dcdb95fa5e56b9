#!/usr/bin/env python
"""Benchmark of the B200-native hot path (contract: one JSON line on stdout from rank 0).

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement on host cores

Headline metric (BASELINE.json): BLS12-381 G1 MSM throughput at 2^22 points, Mpts/s.  A step is one
2^22-point MSM per GPU (device-resident scalars and bases for `value`; host buffers through
tkm_msm_g1_host for `e2e`).  The line also carries the bivariate-NTT figure (16384 x 512, Gelem/s) with its
HBM roofline, the integer-pipe roofline of the MSM's dominant kernel and the CPU baseline.
Multi-GPU: MSM shards by point range, one process per GPU, partial sums gathered with NCCL and combined
on rank 0 (weak scaling: every rank owns 2^22 points).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))

LOG_N = 22
NTT_X, NTT_Y = 16384, 512
NTT_BUTTERFLIES = NTT_X * NTT_Y * ((14 - 2) / 2 + (9 - 2) / 2 + 0.5)  # 83.9 M products per transform
FR_MUL_WIDE_IMADS = 120  # Fr Montgomery product as issued: 64 a*b + 56 reduction wide IMADs (r's lowest limb needs no product; ff.cuh reduce_row)
FR_MUL_WIDE_IMADS_R01 = 112  # the count round 1 was judged on (the p1 product taken as add-with-carry chains instead)
WIDE_PER_MADD = 6 * 288 + 2 * 222 + 432  # XYZZ mixed addition: 6 products + 2 dedicated squarings + the fused two-product Y3
WIDE_PER_AFFINE_ADD = 5 * 288 + 222      # affine pair-tree addition: 3 products of Montgomery's trick + lambda, lambda^2, lambda*(x1 - x3)


def accumulation_work(ctx, entries_fallback):
    """Issued wide IMADs of the last MSM's accumulation phase from the library's own counters (tkm_msm_tree_stats): affine
    additions of the pair-tree levels (n_0 - n_L) plus XYZZ mixed additions of the n_L entries left for the chained pass."""
    L, counts = ctx.msm_tree_stats()
    if L == 0:
        n0 = counts[0] if counts and counts[0] else entries_fallback
        return n0 * WIDE_PER_MADD, {"tree_levels": 0, "mixed_additions": n0}
    tree_adds = counts[0] - counts[L]
    return tree_adds * WIDE_PER_AFFINE_ADD + counts[L] * WIDE_PER_MADD, {"tree_levels": L, "entries_per_level": counts, "affine_additions": tree_adds, "mixed_additions": counts[L]}
METRIC = "BLS12-381 G1 MSM throughput at 2^22 points"
UNIT = "Mpts/s"
G1_GEN = (
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)


def bench_config(log_n, world):
    """The workload description both arms print (the reference arm must describe the same job as ours)."""
    return {"workload": f"G1 MSM 2^{log_n} points per GPU, uniform random scalars (seed 2000+rank), distinct bases k_i*G (k_i uniform, seed 1000+rank)",
            "log2_points_per_gpu": log_n, "parallelism": f"point-range shards x{world}, partial sums all-gathered and combined on rank 0" if world > 1 else "single GPU",
            "l2": "inputs (0.5 GiB) larger than the 126 MB L2"}


# dram__bytes_read.sum + dram__bytes_write.sum from the committed ncu --set full captures of a 2^22-point MSM and a 16384 x 512
# forward transform (profiles/r02_ncu_k_tree_apply_summary.csv, r02_ncu_k_tree_fwd_summary.csv, r02_ncu_k_accumulate_summary.csv,
# r02_ncu_k_ntt_pass_summary.csv).  tree_level0: both halves of k_tree_fwd (4.489 + 4.049 GB) and k_tree_apply (8.713 + 7.866 GB) at
# level 0, 67.1 M digit entries; the deeper levels were not captured and are scaled by their entry counts.
NCU_TRAFFIC = {"tree_level0": 25.117, "tree_level0_entries": 67.1e6, "k_accumulate_tail": 0.552, "k_ntt_pass_x3": 1.464}


def accumulation_traffic_gb(detail):
    """DRAM traffic of one accumulation phase from the level-0 capture: a level's passes move bytes in proportion to its entries."""
    e = detail.get("entries_per_level") or []
    if len(e) < 2:
        return None  # no pair tree in this MSM: nothing captured for the chained form this round
    return NCU_TRAFFIC["tree_level0"] * sum(e[:-1]) / NCU_TRAFFIC["tree_level0_entries"] + NCU_TRAFFIC["k_accumulate_tail"]


def msm_windows(n):
    """Digit windows per scalar the library picks for an n-point MSM (pick_geom in csrc/msm.cu, restated: GLV halves of 128
    bits; cost in wide-IMAD units with the pair tree's cheaper additions, its per-level fixed cost, the sort passes and the
    measured window-reduction cost): (c, W)."""
    best, bc = None, 8
    for c in range(4, 21):
        W = 2 * ((128 + c - 1) // c)
        B = 1 << (c - 1)
        M, avg, L = W * n, n / B, 0
        if avg >= 16 and M >= (1 << 24):
            while (8 << (L + 1)) <= avg and L < 8:
                L += 1
        if L == 0:
            cost = 2604.0 * W * (n + 3.0 * B)
        else:
            R = M / (1 << L)
            cost = (M - R) * 1662.0 + R * 2604.0 + W * B * 18000.0 + ((c + 7) // 8) * M * 75.0 + L * 3.0e9
        if best is None or cost < best:
            best, bc = cost, c
    return bc, 2 * ((128 + bc - 1) // bc)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def random_scalars(seed, n):
    """Uniform 254-bit values (< r), canonical little-endian u64 limbs."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)
    return a


class ClockSampler:
    """Samples SM clock / throttle reasons / power through NVML from a background thread (in-process: an `nvidia-smi -lms`
    loop beside the benchmark was seen to stall kernel launches for tens of ms); started before the warm-up so that samples
    exist for short timed regions, summarised over the [t0, t1] window of the timed region."""

    PERIOD_S = float(os.environ.get("TKM_BENCH_SAMPLER_PERIOD_S", "0.1"))  # developer knob; <= 0 disables sampling

    def __init__(self, index):
        self.index = index
        self.samples = []  # (t, sm_mhz, max_mhz, power_w, reasons_bitmask)
        self.ok = False
        self._stop = threading.Event()

    def start(self):
        if self.PERIOD_S <= 0:
            return
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            self.nv = nv
            self.dev = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.dev, nv.NVML_CLOCK_SM))
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(self.dev) / 1000.0
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev))
                self.samples.append((time.perf_counter(), sm, self.max_mhz, pw, rs))
            except Exception:
                pass
            self._stop.wait(self.PERIOD_S)

    def stop(self):
        self._stop.set()

    def summary(self, t0, t1):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        # NVML clocks-event-reason bits
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        sel = [x for x in self.samples if t0 <= x[0] <= t1 + self.PERIOD_S]
        if not sel and self.samples:  # very short region: take the nearest sample
            sel = [min(self.samples, key=lambda x: abs(x[0] - t1))]
        reasons = sorted({nm for x in sel for nm, b in bits.items() if x[4] & b})
        return {"sm_mhz": float(np.median([x[1] for x in sel])) if sel else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sel), "power_w_max": max(x[3] for x in sel) if sel else None,
                "source": f"NVML, in-process thread, {int(self.PERIOD_S * 1000)} ms period"}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off (sysfs), so that the pinned staging buffers of the e2e leg are
    allocated on that node: eight ranks copying 512 MiB each through one remote node's memory controller is what limited the
    N = 8 e2e line in round 1.  Best effort: returns the node, or None when the topology cannot be read."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU path, restated (oracle/oracle.c: bucket-method MSM on all host threads).
    The reference itself (Rust + un-vendored ICICLE v3.8.0 CPU backend) cannot be built here, so kind = "port"."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_ffi as O

    O.build()
    O.set_num_threads(len(os.sched_getaffinity(0)))  # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1)
    # the SAME workload as rank 0 of our arm: 2^log_n points, the same seeds (bases k_i*G from seed 1000, scalars from seed 2000)
    n = 1 << args.log_n
    G = np.frombuffer(G1_GEN[0].to_bytes(48, "little") + G1_GEN[1].to_bytes(48, "little"), dtype=np.uint64).copy()
    t_gen = time.perf_counter()
    ks = random_scalars(1000, n)
    bases = O.g1_fixed_base_mul_batch(G, ks)  # untimed input generation (our arm does this on the GPU)
    scalars = random_scalars(2000, n)
    t_gen = time.perf_counter() - t_gen
    for _ in range(max(1, min(args.warmup, 1))):
        res = O.msm_g1(scalars, bases)
    assert np.array_equal(res, O.g1_mul(G, O.fr_inner_product(scalars, ks))), "CPU MSM differs from (sum s_i k_i)*G"
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.msm_g1(scalars, bases)
    t_msm = time.perf_counter() - t0
    dt = t_msm / args.steps
    val = n / dt / 1e6
    cores = O.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Fr 255-bit scalars, Fq 381-bit coordinates)",
        "data": "synthetic", "config": bench_config(args.log_n, max(1, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "nproc": os.cpu_count(), "kind": "port",
                         "sample": f"the whole 2^{args.log_n}-point MSM of rank 0 per step (same seeds as the GPU arm; no reduction), oracle/oracle.c Pippenger, OpenMP; "
                                   f"input generation {t_gen:.1f} s untimed; result checked against (sum s_i k_i)*G"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "published_reference_cpu": {"value": 1.01, "unit": UNIT, "note": "ICICLE CPU backend, 8192x511 pts, unnamed macOS host (BASELINE.md)"},
    }
    if not args.skip_prove and args.gpus == 1:
        # "prove s/tx" on the CPU path: the protocol driver on the oracle backend at the FULL reference shape (n=4096, s_max=256,
        # m_I=4096; about 35 s on 16 cores), CRS generated on the CPU first (not timed).  If the MSM steps above already used the
        # time budget of this arm (slow / few host cores) the shape with every extent / 4 is proved instead, and labelled so.
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        sys.path.insert(0, os.path.join(ROOT, "tokamak-zk-evm_b200"))
        import prove_full
        from oracle_backend import OracleBackend
        from tokamak_b200.protocol import synthetic as S

        full = (t_gen + t_msm) < 240.0 and not args.reference_prove_reduced
        cpu = prove_full.run(OracleBackend(), S.reference_shape() if full else prove_full.reduced_shape(), repeats=1, verify=False, from_files=False)
        line["prove"] = {"metric": "prove s/tx", "value": cpu["prove_s"], "unit": "s", "higher_is_better": False, "cores": cores, "kind": "port",
                         "sample": "the full reference shape (n=4096, s_max=256, m_I=4096, l=728): same synthetic circuit, CRS and fixed blinding as the GPU arm's `prove`" if full
                         else "reference shape with every extent / 4 (n=1024, s_max=64, m_I=1024)", "full_shape": full, "setup_s_not_timed": cpu["setup_s"],
                         "stage_s": {k: cpu["median_run"][k] for k in ("init_s", "prove0_s", "prove1_s", "prove2_s", "prove3_s", "prove4_s", "encode_s")},
                         "proof_sha256": cpu["proof_sha256"],
                         "published_reference_cpu": {"value": 45.7, "unit": "s", "note": "full shape, real template tx, ICICLE CPU backend (BASELINE.md)"}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N, help="developer override of the MSM size (the graded config is 22)")
    ap.add_argument("--skip-aux", action="store_true", help="skip the biNTT / cpu-baseline / e2e legs (profiling runs)")
    ap.add_argument("--skip-strong", action="store_true", help="skip the 2^24 strong-scaling MSM leg")
    ap.add_argument("--skip-prove", action="store_true", help="skip the full-prove leg (setup + prove0..4 + verify at the reference shape)")
    ap.add_argument("--reference-prove-reduced", action="store_true", help="reference arm: prove the shape / 4 instead of the full reference shape")
    ap.add_argument("--no-cpu-prove-full", dest="cpu_prove_full", action="store_false",
                    help="skip the CPU oracle's FULL reference-shape prove (about 35 s on 16 cores); the shape / 4 sample always runs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # keep stdout clean for the one JSON line (NCCL / libraries may print banners): route fd 1 to stderr until the end
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import tokamak_b200 as T

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None  # before any pinned allocation: first touch places the staging buffers
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime

        # a short collective timeout: a mismatched collective must fail in minutes, not hold the GPUs for NCCL's default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    ctx = T.Context(local_rank)
    lib, h = ctx.lib, ctx.h
    n = 1 << args.log_n
    warm = max(args.warmup, 3)

    # ---- synthetic inputs: this rank's point range [rank*n, (rank+1)*n): bases k_i*G generated on the device
    G = np.frombuffer(G1_GEN[0].to_bytes(48, "little") + G1_GEN[1].to_bytes(48, "little"), dtype=np.uint64).copy()
    ks = random_scalars(1000 + rank, n)
    scalars = random_scalars(2000 + rank, n)
    d_k = ctx.upload_fr(ks, to_mont=False)
    d_bases = ctx.dev_alloc(n * 96)
    T.check(lib.tkm_g1_fixed_base_mul(h, G.ctypes.data, d_k, 0, n, d_bases))
    # pinned host copies for the e2e leg (canonical bytes, as the reference hands them to msm::msm)
    h_bases = torch.empty((n, 12), dtype=torch.int64, pin_memory=True)
    h_scalars = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    T.check(lib.tkm_memcpy_d2h(h, h_bases.data_ptr(), d_bases, n * 96))
    h_scalars.numpy().view(np.uint64)[:] = scalars
    T.check(lib.tkm_g1_bases_to_mont(h, d_bases, d_bases, n))
    d_scalars = ctx.upload_fr(scalars, to_mont=False)
    ctx.dev_free(d_k)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def combine(partial):
        """Partial sums of all ranks -> total on rank 0 (144-byte-class exchange)."""
        if world == 1:
            return partial
        t = torch.from_numpy(partial.view(np.int64).copy()).cuda()
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        if rank != 0:
            return partial
        return ctx.g1_sum(torch.stack(out).cpu().numpy().view(np.uint64))

    def step_resident():
        return combine(ctx.msm_g1_dev(d_scalars, False, d_bases, n))

    def step_e2e():
        out = np.zeros(12, dtype=np.uint64)
        T.check(lib.tkm_msm_g1_host(h, h_scalars.data_ptr(), h_bases.data_ptr(), n, out.ctypes.data))
        return combine(out)

    def timed(fn, steps):
        barrier()
        l0 = ctx.launch_count()
        ctx.time_begin()
        t0 = time.perf_counter()
        per_step = []
        for _ in range(steps):
            ts = time.perf_counter()
            res = fn()
            per_step.append((time.perf_counter() - ts) * 1e3)  # every step ends with a synchronous 96-byte read-back
        timed.last_per_step = per_step
        ms = ctx.time_end()
        t1 = time.perf_counter()
        wall = (t1 - t0) * 1e3
        barrier()
        clocks = sampler.summary(t0, t1)
        launches = ctx.launch_count() - l0
        # device-event time on the launching stream; the NCCL combine (N>1) runs on torch's stream, so take the larger of
        # the event time and the host wall time around the same region, then the max over ranks
        t = max(ms, wall) if world > 1 else ms
        if world > 1:
            tt = torch.tensor([t], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t / steps, launches, clocks, res

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warm):
        r0 = step_resident()
    ms_res, launches, clocks, r1 = timed(step_resident, args.steps)
    assert np.array_equal(r0, r1), "MSM result is not deterministic"
    value = world * n / ms_res / 1e3  # Mpts/s, whole job

    # ---- the timed result against its expected value at every N: the bases are k_i*G with known k_i, so the combined MSM
    # must equal (sum over ranks of sum_i s_i k_i) * G.  Checker only (CPU oracle: one inner product per rank, one scalar
    # multiplication on rank 0), after the timed region.
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_ffi as O

    O.build()
    O.set_num_threads(max(1, len(os.sched_getaffinity(0)) // max(1, world)))
    R_MOD = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001

    def expected_from_dlogs(sc, kk):
        """(sum_i sc_i kk_i over all ranks) * G on rank 0 (None elsewhere)."""
        ip = O.fr_inner_product(sc, kk)
        if world > 1:
            t = torch.from_numpy(ip.view(np.int64).copy()).cuda()
            outs = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            ips = [o.cpu().numpy().view(np.uint64) for o in outs]
        else:
            ips = [ip]
        if rank != 0:
            return None
        tot = sum(int.from_bytes(v.tobytes(), "little") for v in ips) % R_MOD
        return O.g1_mul(G, np.frombuffer(tot.to_bytes(32, "little"), dtype=np.uint64).copy())

    exp_main = expected_from_dlogs(scalars, ks)
    if rank == 0:
        assert np.array_equal(r1, exp_main), f"timed 2^{args.log_n} x {world} MSM differs from (sum s_i k_i)*G"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_res,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (Fr 255-bit scalars, Fq 381-bit coordinates)",
        "data": "synthetic",
        "config": bench_config(args.log_n, world),
        "checks": {"msm_timed_result": f"combined result of the timed steps == (sum over {world} rank(s) of sum_i s_i k_i)*G (known discrete logs; CPU oracle as checker, outside the timed region)"},
        "clocks": clocks, "gpu_launches": int(launches), "numa_node_rank0": numa,
        "ms_per_step_host_clock": {"min": min(timed.last_per_step), "max": max(timed.last_per_step), "all": [round(x, 3) for x in timed.last_per_step]},
    }

    if not args.skip_aux:
        # ---- e2e: same metric through tkm_msm_g1_host with pinned HOST buffers (H2D + D2H inside the timed region)
        e2e_steps = max(2, min(args.steps, 5))
        step_e2e()
        ms_e2e, _, _, r2 = timed(step_e2e, e2e_steps)
        assert np.array_equal(r1, r2), "host-buffer MSM differs from device-resident MSM"
        line["e2e"] = {"value": world * n / ms_e2e / 1e3, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * (32 + 96), "d2h_bytes_per_step": 96,
                       "api": "tkm_msm_g1_host (replaces msm::msm with HostSlice scalars and bases)"}

        if rank == 0:
            # ---- integer-pipe roofline of the dominant kernel (k_accumulate): SURVEY.md §8d
            imad_wide = ctx.microbench(1)
            imad_wide_x = ctx.microbench(5)
            fq_mul = ctx.microbench(3)
            madd = ctx.microbench(4)
            W = msm_windows(n)[1]
            bfly = ctx.microbench(6)
            line["microbench"] = {"imad_wide_u32_per_s": imad_wide, "imad_wide_x_u32_per_s": imad_wide_x, "imad_u32_per_s": ctx.microbench(0),
                                  "fr_mul_per_s": ctx.microbench(2), "fq_mul_per_s": fq_mul, "xyzz_madd_per_s": madd, "fr_butterfly_per_s": bfly}
            if W:
                # k_accumulate alone, timed live with CUDA events on the launching stream (tkm_kernel_time_last), averaged
                acc_ms = []
                for _ in range(max(3, min(args.steps, 10))):
                    ctx.msm_g1_dev(d_scalars, False, d_bases, n)  # local MSM only: this leg runs on rank 0 alone, no collective
                    acc_ms.append(ctx.kernel_time_last())
                acc_ms = float(np.mean(acc_ms))
                adds = n * W
                # wide IMADs the accumulation phase actually issues, from the library's per-level entry counts: affine pair-tree
                # additions x 1662 (5 products x 288 + 1 squaring x 222) + the chained XYZZ mixed additions of the remaining
                # entries x 2604 (6 x 288 + 2 x 222 + the fused two-product Y3 x 432); SURVEY 8d's nominal count is 10 x 288 per addition
                work, detail = accumulation_work(ctx, adds)
                ach = work / (acc_ms * 1e-3)
                peak = max(imad_wide, imad_wide_x)
                line["roofline"] = {"bound": "int32", "achieved": ach / 1e12, "peak": peak / 1e12, "unit": "T(32x32+64 IMAD.WIDE)/s", "frac": ach / peak,
                                    "traffic": accumulation_traffic_gb(detail), "traffic_unit": "GB per accumulation phase (dram read+write: level-0 tree launches and the XYZZ pass from the ncu --set full "
                                    "captures in profiles/, deeper levels scaled by their entry counts); the pair tree trades bytes for multiplications: ~400 B per addition against 104 B for the chained form",
                                    "algorithmic_gb": adds * 104 / 1e9, "kernel": "accumulation phase: k_tree_fwd/k_tree_apply per level + k_accumulate", "kernel_ms": acc_ms,
                                    "kernel_share_of_step": acc_ms / ms_res, "work": detail,
                                    "note": "MSM is integer-pipe bound (SURVEY.md 8d; the schema's hbm/tensor bounds do not describe it): issued wide IMADs of the "
                                            "bucket accumulation (affine pair-tree additions x 1662 + chained XYZZ mixed additions x 2604, counts from tkm_msm_tree_stats) "
                                            "/ the phase's event-timed duration; peak = best measured IMAD.WIDE.U32 stream on this GPU (64-bit-addend or carry-chain form)",
                                    "whole_msm_frac": work / (ms_res * 1e-3) / peak,
                                    "useful_work_frac_vs_xyzz_only": adds * WIDE_PER_MADD / (acc_ms * 1e-3) / peak,
                                    # the schema's HBM view of the same kernel, for completeness: 104 algorithmic bytes per addition
                                    "hbm_view": {"bound": "hbm", "achieved": adds * 104 / (acc_ms * 1e-3) / 1e9, "peak": measured_peaks()[0], "unit": "GB/s",
                                                 "frac": adds * 104 / (acc_ms * 1e-3) / 1e9 / measured_peaks()[0]},
                                    "madd_frac": adds / (acc_ms * 1e-3) / madd}

            # ---- bivariate NTT 16384 x 512 (device-resident) with its HBM roofline
            hbm, how = measured_peaks()
            nn = NTT_X * NTT_Y
            ctx.init_ntt_domain_for_size(nn)
            d_poly = ctx.upload_fr(random_scalars(3000, nn), to_mont=False)
            for direction, key in ((T.FORWARD, "forward"), (T.INVERSE, "inverse")):
                for _ in range(warm):
                    ctx.bintt_dev(d_poly, d_poly, NTT_X, NTT_Y, direction)
                ctx.sync()
                l0 = ctx.launch_count()
                ctx.time_begin()
                for _ in range(args.steps):
                    ctx.bintt_dev(d_poly, d_poly, NTT_X, NTT_Y, direction)
                ms = ctx.time_end() / args.steps
                kms = ctx.kernel_time_last()  # the k_ntt_pass launches of the last transform alone
                ach = 128.0 * nn / (ms * 1e-3) / 1e9
                line.setdefault("bintt", {})[key] = {
                    "shape": [NTT_X, NTT_Y], "ms": ms, "value": nn / ms / 1e6, "unit": "Gelem/s", "launches_per_transform": (ctx.launch_count() - l0) // args.steps,
                    "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": NCU_TRAFFIC["k_ntt_pass_x3"],
                                 "traffic_unit": "GB per transform (dram read+write summed over the 3 launches, ncu --set full capture in profiles/)",
                                 "peak_source": f"{how} (MEASURED_PEAKS.json hbm_gbs)", "algorithmic_bytes_per_element": 128,
                                 "kernel": "k_ntt_pass x3", "kernel_ms": kms},
                    # the bound that actually binds: 256-bit modular butterflies on the INT32 pipes
                    "roofline_int32": {"bound": "int32", "unit": "T(32x32+64 IMAD.WIDE)/s", "achieved": NTT_BUTTERFLIES * FR_MUL_WIDE_IMADS / (kms * 1e-3) / 1e12,
                                       "peak": max(imad_wide, imad_wide_x) / 1e12,
                                       "frac": NTT_BUTTERFLIES * FR_MUL_WIDE_IMADS / (kms * 1e-3) / max(imad_wide, imad_wide_x),
                                       "frac_at_112_per_product": NTT_BUTTERFLIES * FR_MUL_WIDE_IMADS_R01 / (kms * 1e-3) / max(imad_wide, imad_wide_x),
                                       "butterfly_stream": {"achieved_g_per_s": NTT_BUTTERFLIES / (kms * 1e-3) / 1e9, "measured_stream_g_per_s": bfly / 1e9},
                                       "note": "product-carrying butterflies of a 16384x512 transform: N*((log2 x - 2)/2 + (log2 y - 2)/2 + 1/2) = 83.9 M "
                                               "(the radix-4 tail of each axis needs one product per four elements) x 120 wide IMADs issued per Fr product "
                                               "(64 a*b + 56 reduction as fused carry chains; ff.cuh; frac_at_112_per_product restates it on round 1's count) / the k_ntt_pass launches' event-timed duration; peak = the same measured "
                                               "IMAD.WIDE.U32 issue peak the MSM roofline uses"}}
            # e2e biNTT through the host-buffer entry point (pinned buffers)
            h_poly = torch.empty((nn, 4), dtype=torch.int64, pin_memory=True)
            h_out = torch.empty((nn, 4), dtype=torch.int64, pin_memory=True)
            h_poly.numpy().view(np.uint64)[:] = random_scalars(3001, nn)
            T.check(lib.tkm_bintt_host(h, h_poly.data_ptr(), h_out.data_ptr(), NTT_X, NTT_Y, T.FORWARD, None, None))
            ctx.time_begin()
            for _ in range(3):
                T.check(lib.tkm_bintt_host(h, h_poly.data_ptr(), h_out.data_ptr(), NTT_X, NTT_Y, T.FORWARD, None, None))
            ms = ctx.time_end() / 3
            line["bintt"]["e2e_forward"] = {"ms": ms, "value": nn / ms / 1e6, "unit": "Gelem/s", "h2d_bytes_per_step": nn * 32, "d2h_bytes_per_step": nn * 32,
                                            "api": "tkm_bintt_host (replaces _biNTT with HostSlice in/out)"}
            ctx.dev_free(d_poly)

        if rank == 0:
            # ---- polynomial-engine kernels (HBM-bound: a few field operations per 32-byte element): achieved GB/s of each against
            # the measured HBM peak.  Event-timed over the C-ABI call on device-resident polynomials at the prover's shapes;
            # algorithmic bytes = operands read once + result written once.
            hbm_p, _ = measured_peaks()
            px, py = 8192, 512
            npts = px * py
            ctx.init_ntt_domain_for_size(NTT_X * NTT_Y)
            mk = lambda seed, x_, y_: T.DensePolynomialExt.from_coeffs(ctx, random_scalars(seed, x_ * y_), x_, y_)
            pa, pb, pc = mk(7001, px, py), mk(7002, px, py), mk(7003, px, py)
            sc = [0x1234567 + (1 << 200), 0x7654321 + (1 << 199), 0xABCDEF + (1 << 198)]
            eng = {}

            fr_mul_peak = line.get("microbench", {}).get("fr_mul_per_s")

            def timed_op(name, fn, bytes_moved, reps=5, note=None, muls_per_elem=0, elems=0):
                fn()
                ctx.sync()
                l0_ = ctx.launch_count()
                ctx.time_begin()
                for _ in range(reps):
                    keep = fn()
                ms_ = ctx.time_end() / reps
                eng[name] = {"ms": ms_, "algorithmic_gb": bytes_moved / 1e9, "achieved_gbs": bytes_moved / (ms_ * 1e-3) / 1e9,
                             "frac_of_hbm": bytes_moved / (ms_ * 1e-3) / 1e9 / hbm_p, "launches": (ctx.launch_count() - l0_) // reps}
                if muls_per_elem and fr_mul_peak:  # kernels with several products per element are bound by the Fr multiplier, not HBM
                    eng[name]["fr_mul_frac"] = muls_per_elem * elems / (ms_ * 1e-3) / fr_mul_peak
                    eng[name]["bound"] = "Fr multiplier (int32 pipe)" if eng[name]["fr_mul_frac"] > eng[name]["frac_of_hbm"] else "hbm"
                if note:
                    eng[name]["note"] = note
                del keep

            timed_op("k_lincomb (poly_comb!, 3 terms 8192x512)", lambda: T.DensePolynomialExt.lincomb([(sc[0], pa), (sc[1], pb), (sc[2], pc)]), 4 * npts * 32,
                     muls_per_elem=3, elems=npts)
            timed_op("k_lincomb (shifted helper: c0 p + c1 X p)", lambda: T.DensePolynomialExt.lincomb([(sc[0], pa, 0, 0), (sc[1], pa, 1, 0)]), (1 + 2) * npts * 32,
                     note="one operand read (twice, second time from L2), result 16384x512", muls_per_elem=2, elems=npts)
            timed_op("k_axpby (p + q)", lambda: pa + pb, 3 * npts * 32)
            timed_op("k_scale_coeffs (scale_coeffs_x)", lambda: pa.scale_coeffs_x(sc[0]), 2 * npts * 32, muls_per_elem=1, elems=npts)
            timed_op("k_eval_partial + k_sum_partials (eval at a point)", lambda: pa.eval(sc[0], sc[1]), npts * 32)
            timed_op("k_vanish_qy/_qx (div_by_vanishing_opt c=4096 d=256)", lambda: pa.clone().div_by_vanishing_opt(4096, 256), (1 + 2 + 1 + 1) * npts * 32,
                     note="includes the clone (the reference's &mut self): read+write clone, read, write Q_X, Q_Y/B traffic ~ N")
            timed_op("k_ruffini_seg_* + k_ruffini_y_scan (div_by_ruffini)", lambda: pa.div_by_ruffini(sc[0], sc[1]), 2 * npts * 32)
            d_t0, d_t1 = ctx.dev_alloc(npts * 32), ctx.dev_alloc(npts * 32)
            T.check(lib.tkm_fr_vec_fill(h, T.fr_bytes(5)[1], d_t0, npts))
            timed_op("k_transpose (4096 x 1024)", lambda: ctx.transpose_dev(d_t0, d_t1, 4096, 1024), 2 * npts * 32)
            timed_op("k_vec_op (pointwise mul)", lambda: T.check(lib.tkm_fr_vec_op(h, 2, d_t0, d_t1, d_t1, npts)), 3 * npts * 32)
            ctx.dev_free(d_t0)
            ctx.dev_free(d_t1)
            # the fused expression kernel on prove2's p_comb (7 leaves, domain 16384 x 512): kernel-only time from its own events
            m_i, s_mx = 4096, 256
            lv = [mk(7100 + k, m_i, s_mx // 2) for k in range(7)]  # triple products must fit the 16384 x 512 domain
            r_, g_, f_, r1_, r2_, KL_, K0_ = lv
            E = T.PolyExpr
            rg = E.mul(E.poly(r_), E.poly(g_))
            expr = E.weighted_sum([(1, E.mul(E.sub(E.poly(r_), E.scalar(1)), E.poly(KL_))),
                                   (sc[0], E.mul_x_minus_one(E.sub(rg, E.mul(E.poly(r1_), E.poly(f_))))),
                                   (sc[1], E.mul(E.poly(K0_), E.sub(rg, E.mul(E.poly(r2_), E.poly(f_)))))])
            for _ in range(2):
                res_ = expr.evaluate_fused_with_domain(NTT_X, NTT_Y)
            ctx.sync()
            t0_ = time.perf_counter()
            res_ = expr.evaluate_fused_with_domain(NTT_X, NTT_Y)
            ctx.sync()
            whole_ms = (time.perf_counter() - t0_) * 1e3
            kms_ = ctx.poly_kernel_time_last()
            bytes_ = (7 + 1) * NTT_X * NTT_Y * 32
            eng["k_polyexpr (p_comb, 7 leaves, 16384x512)"] = {"ms": kms_, "algorithmic_gb": bytes_ / 1e9, "achieved_gbs": bytes_ / (kms_ * 1e-3) / 1e9,
                                                              "frac_of_hbm": bytes_ / (kms_ * 1e-3) / 1e9 / hbm_p, "launches": 1,
                                                              "whole_evaluate_fused_ms": whole_ms, "bound": "Fr multiplier (int32 pipe)",
                                                              "fr_mul_frac": (8 * NTT_X * NTT_Y / (kms_ * 1e-3) / fr_mul_peak) if fr_mul_peak else None,
                                                              "note": "8 Fr products + 6 additions per element next to 256 B of traffic: the pointwise DAG of "
                                                                      "PolyExpr::evaluate_on_domain as ONE kernel (the reference: ~12 passes with a fresh 256 MiB vector each); "
                                                                      "whole call = 7 leaf pads + 7 forward biNTTs + this kernel + 1 inverse biNTT"}
            del res_, lv, pa, pb, pc
            line["poly_engine"] = {"peak_gbs": hbm_p, "kernels": eng}

        if rank == 0 and world == 1:
            # ---- CPU baseline beside it (N=1 only): the oracle port on the host cores, bounded sample, also the bit-exact check
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle_ffi as O

            O.build()
            O.set_num_threads(len(os.sched_getaffinity(0)))
            ls = min(18, args.log_n)
            ns = 1 << ls
            hb = h_bases.numpy().view(np.uint64)[:ns]
            hs = h_scalars.numpy().view(np.uint64)[:ns]
            t0 = time.perf_counter()
            exp = O.msm_g1(hs, hb)
            dt = time.perf_counter() - t0
            got = ctx.msm_g1_host(hs, hb)
            assert np.array_equal(got, exp), "GPU MSM differs from the CPU oracle on the sample"
            x2, y2 = 4096, 256
            a = O.random_fr(7, x2 * y2)
            t0 = time.perf_counter()
            ev = O.bintt(a, x2, y2, False)
            dt_ntt = time.perf_counter() - t0
            assert np.array_equal(ctx.bintt_host(a, x2, y2, T.FORWARD), ev), "GPU biNTT differs from the CPU oracle on the sample"
            line["cpu_baseline"] = {"value": ns / dt / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
                                    "sample": f"first 2^{ls} points of the same inputs, oracle/oracle.c Pippenger (OpenMP); result compared bit-exactly with the GPU",
                                    "bintt": {"value": x2 * y2 / dt_ntt / 1e9, "unit": "Gelem/s", "sample": "4096x256 forward biNTT, oracle/oracle.c radix-2 (OpenMP)"},
                                    "published_reference": "ICICLE CPU backend: 1.01 Mpts/s at 8192x511 pts; biNTT 2^23 forward 497 ms (unnamed macOS host, BASELINE.md)"}

    if not args.skip_aux and not args.skip_strong:
        # ---- strong scaling (north star): ONE 2^24-point MSM cut into point ranges over the N ranks (2^24 / N points each),
        # partial sums combined on rank 0; result checked against the known discrete logs; k_accumulate's roofline per N
        LOG_S = 24
        ns = (1 << LOG_S) // world
        ks2, sc2 = random_scalars(5000 + rank, ns), random_scalars(6000 + rank, ns)
        d_k2 = ctx.upload_fr(ks2, to_mont=False)
        d_b2 = ctx.dev_alloc(ns * 96)
        T.check(lib.tkm_g1_fixed_base_mul(h, G.ctypes.data, d_k2, 0, ns, d_b2))
        T.check(lib.tkm_g1_bases_to_mont(h, d_b2, d_b2, ns))
        ctx.dev_free(d_k2)
        d_s2 = ctx.upload_fr(sc2, to_mont=False)

        def step_strong():
            return combine(ctx.msm_g1_dev(d_s2, False, d_b2, ns))

        for _ in range(3):
            step_strong()
        ms_strong, _, clocks_s, rs = timed(step_strong, max(3, min(args.steps, 10)))
        acc_s = ctx.kernel_time_last()
        work_s, detail_s = accumulation_work(ctx, ns * msm_windows(ns)[1])  # this rank's counts; ranks differ by well under 1 %
        if world > 1:
            tt = torch.tensor([acc_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            acc_s = float(tt.item())
        exp_s = expected_from_dlogs(sc2, ks2)
        for p_ in (d_b2, d_s2):
            ctx.dev_free(p_)
        if rank == 0:
            assert np.array_equal(rs, exp_s), f"2^{LOG_S} MSM over {world} rank(s) differs from (sum s_i k_i)*G"
            c_s, W_s = msm_windows(ns)
            peak_s = line.get("roofline", {}).get("peak")
            work = work_s
            line["msm_2p24_strong"] = {
                "metric": "one 2^24-point G1 MSM, point ranges over N GPUs", "log2_points_total": LOG_S, "ranks": world, "points_per_rank": ns,
                "ms": ms_strong, "value": (1 << LOG_S) / ms_strong / 1e3, "unit": UNIT, "scaling": "strong", "window_bits": c_s, "windows": W_s,
                "checked": "combined result == (sum s_i k_i)*G over all ranks (known discrete logs)",
                "roofline": {"bound": "int32", "kernel": "accumulation phase (pair tree + k_accumulate)", "kernel_ms_max_over_ranks": acc_s, "work": detail_s, "achieved": work / (acc_s * 1e-3) / 1e12,
                             "peak": peak_s, "unit": "T(32x32+64 IMAD.WIDE)/s per GPU", "frac": (work / (acc_s * 1e-3) / 1e12 / peak_s) if peak_s else None,
                             "whole_msm_frac": (work / (ms_strong * 1e-3) / 1e12 / peak_s) if peak_s else None},
                "clocks": clocks_s}
    if rank == 0 and world == 1 and not args.skip_aux and not args.skip_prove:
        # ---- "prove s/tx" (first part of BASELINE.json's metric): full Prover.init + prove0..prove4 on this GPU at the
        # reference's circuit shape, proof checked by the restated verifier; then the CPU baseline beside it on a bounded
        # sample (every extent / 4), where the GPU and CPU proofs must be byte-identical
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import copy

        import prove_full
        from tokamak_b200.protocol import synthetic as S
        from tokamak_b200.protocol.backend import GpuBackend

        for p_ in (d_bases, d_scalars):
            ctx.dev_free(p_)
        gpu_be = GpuBackend(ctx)
        full = prove_full.run(gpu_be, S.reference_shape(), repeats=3, verify=True, fixed_base_tables=True, sync=ctx.sync, warmup=3, crs_load=True,
                              keep_sigma=args.cpu_prove_full)
        full_sigma = full.pop("_sigma", None)
        full["reference"] = {"cpu_prove_s": 45.7, "icicle_cuda_prove_s": 21.08, "stage_split_cpu_s": [5.21, 10.09, 2.13, 13.37, 1.56, 13.33],
                             "stage_split_icicle_cuda_s": [0.72, 4.03, 0.78, 7.27, 0.90, 7.37],
                             "source": "BASELINE.md (reference's own artifacts, unnamed hosts, real template tx, 166 placements)"}
        line["prove"] = {"metric": "prove s/tx", "value": full["prove_s"], "unit": "s", "higher_is_better": False, **full}
        # ---- the reference's own circuit library (14 subcircuits, real constraint sparsity, m_D = 26591) with 166 placements like
        # the template transaction; the witness is seeded and does NOT satisfy the constraints (the circom witness calculators
        # cannot run here), so this leg is timing-only: same kernels, same sizes, a proof nobody should accept
        real_path = os.path.join(ROOT, "tests", "golden", "real_library.json.xz")
        if os.path.exists(real_path):
            real = prove_full.run(gpu_be, None, repeats=3, verify=False, sync=ctx.sync, warmup=2, from_files=False,
                                  inputs=prove_full.real_library_inputs(real_path))
            line["prove"]["real_library"] = {
                "prove_s": real["prove_s"], "median_run": real["median_run"], "shape": real["shape"], "setup_s": real["setup_s"],
                "note": "packages/frontend/qap-compiler/subcircuits/library (packed fixture tests/golden/real_library.json.xz), 166 placements, seeded "
                        "unsatisfying witness: timing only"}
        small = prove_full.run(gpu_be, prove_full.reduced_shape(), repeats=3, verify=False, sync=ctx.sync, keep_sigma=True, warmup=1)
        from oracle_backend import OracleBackend, OracleTable

        sg = copy.copy(small.pop("_sigma"))
        for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
            t_ = getattr(sg, name)
            setattr(sg, name, OracleTable(t_.points_host(), t_.rows, t_.cols))
        cpu = prove_full.run(OracleBackend(), prove_full.reduced_shape(), repeats=1, verify=False, sigma=sg)
        assert cpu["proof_sha256"] == small["proof_sha256"], "GPU proof and CPU-oracle proof differ on the reduced shape"
        if args.cpu_prove_full:
            sgf = copy.copy(full_sigma)
            for name in ("xy_powers", "gamma_inv_o_inst", "eta_inv_li_o_inter_alpha4_kj", "delta_inv_li_o_prv"):
                t_ = getattr(sgf, name)
                setattr(sgf, name, OracleTable(t_.points_host(), t_.rows, t_.cols))
            cpu_full = prove_full.run(OracleBackend(), S.reference_shape(), repeats=1, verify=False, sigma=sgf)
            assert cpu_full["proof_sha256"] == full["proof_sha256"], "GPU proof and CPU-oracle proof differ at the full reference shape"
            line.setdefault("cpu_baseline", {})["prove_full_shape"] = {
                "value": cpu_full["prove_s"], "unit": "s", "cores": O.num_threads(), "kind": "port", "gpu_same_input_s": full["prove_s"],
                "cpu_stage_s": {k: cpu_full["median_run"][k] for k in ("init_s", "prove0_s", "prove1_s", "prove2_s", "prove3_s", "prove4_s", "encode_s")},
                "sample": "the full reference-shape circuit (no reduction); proof byte-identical to the GPU proof"}
        line.setdefault("cpu_baseline", {})["prove"] = {
            "value": cpu["prove_s"], "unit": "s", "cores": O.num_threads(), "kind": "port",
            "sample": "reference shape with every extent / 4 (n=1024, s_max=64, m_I=1024; 1/16 of the MSM and NTT work): protocol driver on "
                      "the oracle backend (oracle/oracle.c, OpenMP); proof byte-identical to the GPU proof of the same input",
            "gpu_same_sample_s": small["prove_s"], "cpu_stage_s": {k: cpu["median_run"][k] for k in ("init_s", "prove0_s", "prove1_s", "prove2_s", "prove3_s", "prove4_s", "encode_s")}}
    if world > 1 and not args.skip_aux:
        # ---- row-sharded bivariate NTT with the X<->Y transpose as an NCCL all-to-all
        from tokamak_b200 import dist as D

        ops = D.CudaLocalOps(ctx)
        ctx.init_ntt_domain_for_size(NTT_X * NTT_Y)
        nn = NTT_X * NTT_Y
        lo, hi = D.shard_range(NTT_X, world, rank)
        src = torch.from_numpy(random_scalars(4000 + rank, (hi - lo) * NTT_Y).view(np.int64)).cuda().view(hi - lo, NTT_Y, 4)
        # expected evaluations of this rank's column shard from the CPU oracle (checker, outside the timed regions).  The device
        # buffers hold the raw limbs as Montgomery representations; the transform is linear, so the oracle applied to the
        # same raw values (all < r) gives the same raw output.
        O.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))
        full_in = np.concatenate([random_scalars(4000 + r_, (D.shard_range(NTT_X, world, r_)[1] - D.shard_range(NTT_X, world, r_)[0]) * NTT_Y) for r_ in range(world)])
        yb = NTT_Y // world
        exp_cols = torch.from_numpy(np.ascontiguousarray(O.bintt(full_in, NTT_X, NTT_Y, False).reshape(NTT_X, NTT_Y, 4)[:, rank * yb:(rank + 1) * yb]).view(np.int64))
        del full_in
        res = {}
        for key in ("forward", "roundtrip"):
            def one(buf):
                ev = D.bintt_sharded_forward(ops, buf, NTT_X, NTT_Y)
                if key == "roundtrip":
                    ev = D.bintt_sharded_inverse(ops, ev.contiguous(), NTT_X, NTT_Y)
                return ev
            bufs = [src.clone() for _ in range(warm + args.steps)]
            for i in range(warm):
                out = one(bufs[i])
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                out = one(bufs[warm + i])
            e1.record()
            torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[key] = float(tt.item())
            if key == "roundtrip":
                assert torch.equal(out.view(-1), src.view(-1)), "sharded biNTT round trip is not the identity"
            else:
                assert torch.equal(out.reshape(NTT_X, yb, 4).cpu(), exp_cols), "sharded biNTT (NCCL all-to-all) differs from the CPU oracle on this rank's column shard"
            del bufs
        # the same transform with the exchange fused into the last NTT pass (P2P stores over NVLink, no NCCL on the data path)
        fused = None
        if world & (world - 1) == 0:
            try:
                ex = D.PeerExchange(NTT_X, NTT_Y, torch.device("cuda", local_rank))
                fres = {}
                for key in ("forward", "roundtrip"):
                    def one_fused():
                        ev = D.bintt_sharded_forward_fused(ops, ex, src)
                        if key == "roundtrip":
                            return D.bintt_sharded_inverse_fused(ops, ex, ev)
                        return ev
                    for i in range(warm):
                        out = one_fused()
                    barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for i in range(args.steps):
                        out = one_fused()
                    e1.record()
                    torch.cuda.synchronize()
                    tt = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    fres[key] = float(tt.item())
                    if key == "roundtrip":
                        assert torch.equal(out.reshape(-1), src.reshape(-1)), "fused sharded biNTT round trip is not the identity"
                    else:
                        assert torch.equal(out.reshape(NTT_X, yb, 4).cpu(), exp_cols), "fused sharded biNTT differs from the CPU oracle on this rank's column shard"
                fused = {"forward_ms": fres["forward"], "forward_gelem_s": nn / fres["forward"] / 1e6, "roundtrip_ms": fres["roundtrip"],
                         "exchange": "fused into the last k_ntt_pass: 128-bit P2P stores into peer-mapped symmetric memory + 2 stream-ordered barriers",
                         "p2p_bytes_per_rank": (world - 1) * (nn // world // world) * 32}
            except AssertionError:
                raise  # a wrong result is never reported as "unavailable"
            except Exception as exc:  # symmetric memory not available on this box: keep the NCCL numbers, say why
                fused = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        if rank == 0:
            line["bintt_sharded"] = {"shape": [NTT_X, NTT_Y], "ranks": world, "forward_ms": res["forward"], "forward_gelem_s": nn / res["forward"] / 1e6,
                                     "roundtrip_ms": res["roundtrip"], "scaling": "strong", "exchange": "all_to_all_single (NCCL), one per direction",
                                     "alltoall_bytes_per_rank": (world - 1) * (nn // world // world) * 32, "fused_exchange": fused,
                                     "checked": "every rank's column shard of the forward result (NCCL and fused) == CPU oracle biNTT of the gathered input; round trips == input",
                                     "roofline": {"bound": "hbm", "unit": "GB/s per GPU", "peak": measured_peaks()[0],
                                                  "achieved_nccl": 128.0 * nn / world / (res["forward"] * 1e-3) / 1e9,
                                                  "frac_nccl": 128.0 * nn / world / (res["forward"] * 1e-3) / 1e9 / measured_peaks()[0],
                                                  "achieved_fused": (128.0 * nn / world / (fused["forward_ms"] * 1e-3) / 1e9) if fused and "forward_ms" in fused else None,
                                                  "frac_fused": (128.0 * nn / world / (fused["forward_ms"] * 1e-3) / 1e9 / measured_peaks()[0]) if fused and "forward_ms" in fused else None}}
    sampler.stop()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
