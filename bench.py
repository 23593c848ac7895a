#!/usr/bin/env python
"""bench.py — throughput of the SuperDendrix significance path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): the BRAF-shape matrix (synthetic DepMap-20Q2
shape, 770 cell lines x ~400 OncoKB-like features), a fixed k=3 module, and
100 000 Curveball nulls per GPU per step, each generated (5*min(S,F) trades, one
round from a burnt-in pool state) and scored with the SuperDendrix set objective,
exceedances counted for the permutation p-value.  One step = one such pass = one
launch of k_trade_lanes.  Ranks shard the null index range (weak scaling: 100 000
nulls per rank per step); the only exchange is the all-reduce of the counts.

Prints ONE JSON line on rank 0.  `value` = nulls/s with matrix, weights and pool
resident in HBM; `e2e` = the same call sequence a user makes through the C ABI from
host buffers (upload matrix + weights, rebuild the pool, run, read counts) every step.
`roofline` = the popc fraction of SURVEY 8d against the rate measured on the box.
`strong_scaling`: configs[1] as stated — 100 000 nulls IN TOTAL over the N GPUs through
superdendrix_b200.dist.null_pvalue, collectives inside the timed region.
`full_size_configs`: configs[2..4] at full size through superdendrix_b200.dist.
`extra` (N = 1): secondary figures with the CPU port beside them.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from superdendrix_b200 import defaults      # noqa: E402

NULLS_PER_STEP = 100_000
# every null is one ordinary round away from a burnt-in pool state; blocks of 32 chains share their row pairs
CHAIN_LEN, BURN_ROUNDS, POOL_SIZE, PAIR_GROUP = defaults.CHAIN_LEN, defaults.BURN_ROUNDS, defaults.POOL_SIZE, defaults.PAIR_GROUP
if os.environ.get("SDX_BENCH_CHAINS"):                # experiments only: "chain_len,burn,pool,pair_group"
    CHAIN_LEN, BURN_ROUNDS, POOL_SIZE, PAIR_GROUP = (int(x) for x in os.environ["SDX_BENCH_CHAINS"].split(","))
SEED = 1764          # the reference's default -rs (src/generate_null_matrices.py:32)
POPC_PEAK_NOMINAL = 148 * 16 * 1.965e9     # 32-bit popc/s
# profiles/r02_k_trade_lanes_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one 100 000-null launch of this configuration
NCU_DRAM_BYTES_PER_NULL = (7.252417e9 + 8.019517e9) / 100000


def workload():
    from superdendrix_b200 import synth
    dense, samples, features = synth.feature_matrix(770, 400, seed=20)
    w = synth.weights(dense.shape[0], 1, seed=3)
    module = synth.pick_module(dense, 3)
    return dense, w, module


# ---------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.stop_flag = [], set(), threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:        # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:     # noqa: BLE001
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:         # noqa: BLE001
                pass
            time.sleep(self.period)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------
def cpu_nulls_per_second(dense, w, module, n_nulls, seed):
    """The reference's Python path (oracle/refport.py, CPython random, adjacency lists)
    on one core: n_nulls chained Curveball nulls, each scored with SuperW."""
    import random

    from oracle import refport
    S, F = dense.shape
    mat = dense.astype(np.float64)
    samples = list(range(S))
    profile = {s: float(w[0, s]) for s in samples}
    random.seed(seed)
    lists = refport.presences(mat)
    t0 = time.perf_counter()
    ge = 0
    cases = [set(np.nonzero(dense[:, g])[0].tolist()) for g in module]
    z_obs = refport.super_w(cases, profile)
    for _ in range(n_nulls):
        null = refport.curveball(mat, lists)
        cases = [set(np.nonzero(null[:, g])[0].tolist()) for g in module]
        ge += refport.super_w(cases, profile) >= z_obs
    dt = time.perf_counter() - t0
    return n_nulls / dt, dt, ge


def _cpu_worker(args):
    dense, w, module, n, seed = args
    return cpu_nulls_per_second(dense, w, module, n, seed)


def run_reference(args):
    """--impl reference: the reference's CPU path (port; the reference itself cannot travel to
    the GPU box) on all host cores, sharded the way the reference shards: one generator
    process per core with its own seed / null-index prefix (src/generate_null_matrices.py:29,32,82)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    dense, w, module = workload()
    cores = os.cpu_count() or 1
    per_proc = 6
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for step in range(args.warmup + args.steps):
            jobs = [(dense, w, module, per_proc, SEED + 1000 * step + i) for i in range(cores)]
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = cores * per_proc * len(times) / total
    line = {
        "impl": "reference", "metric": "curveball_nulls_scored_per_s", "value": value, "unit": "nulls/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
        "config": {"workload": "C1: BRAF-shape %dx%d (synthetic DepMap-20Q2 shape), fixed k=3 module, %d Curveball nulls per GPU per step, "
                               "%d trades each, fused SuperW scoring + exceedance count" % (dense.shape[0], dense.shape[1], NULLS_PER_STEP,
                                                                                            5 * min(dense.shape)),
                   "sample": "%d nulls per step (%d per process x %d processes) of the 100000-null workload" % (cores * per_proc, per_proc, cores)},
        "cpu_baseline": {"value": value, "unit": "nulls/s", "cores": cores, "kind": "port",
                         "sample": "%d nulls per step, %d processes, oracle/refport.py (CPython %s)" % (cores * per_proc, cores, sys.version.split()[0])},
        "e2e": {"value": value, "unit": "nulls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from superdendrix_b200 import bitpack
    from superdendrix_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsdx has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dense, w, module = workload()
    S, F = dense.shape
    feat_rows = bitpack.pack_rows(dense.T)
    # pinned host staging for the e2e leg
    pin_rows = torch.from_numpy(feat_rows.view(np.int32)).pin_memory()
    pin_w = torch.from_numpy(w).pin_memory()
    rows_np = pin_rows.numpy().view(np.uint32)
    w_np = pin_w.numpy()
    modules = module[None, :].astype(np.int32)

    eng = Engine(local)
    stream = torch.cuda.Stream()          # every launch of the step (ours and torch's flush) goes on this stream
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    eng.upload_matrix(rows_np, S)
    eng.upload_weights(w_np)
    eng.set_burn_in(BURN_ROUNDS, POOL_SIZE)
    eng.set_pair_group(PAIR_GROUP)
    R, L, wpr, ras = eng.trade_shape()
    T = 5 * min(S, F)
    n = args.nulls
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def null_range(step):
        # global null index space: step-major, then rank — disjoint ranges, as the reference's -pre sharding
        return (step * world + rank) * n

    z_obs = eng.null_pvalue(SEED, 0, 0, modules)["z_obs"]
    counts = torch.zeros(1, dtype=torch.int64, device="cuda")

    def step_resident(step):
        r = eng.null_pvalue(SEED, null_range(step), n, modules, chain_len=CHAIN_LEN, z_obs=z_obs)
        return r["n_ge"]

    def step_e2e(step):
        eng.upload_matrix(rows_np, S)
        eng.upload_weights(w_np)
        r = eng.null_pvalue(SEED, null_range(step), n, modules, chain_len=CHAIN_LEN)      # new matrix: the pool is rebuilt
        return r["n_ge"]

    def timed(fn, steps, warmup, first_step):
        for i in range(warmup):
            fn(first_step + i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        l0 = eng.total_launches()
        kernel_ms, plan_ms, ge, tl = 0.0, 0.0, 0, 0
        t0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()                      # L2 flush between timed iterations (256 MiB > 126 MB L2)
            ev[i][0].record(stream)
            ge += int(fn(first_step + warmup + i)[0])
            ev[i][1].record(stream)
            tm = eng.timing()
            kernel_ms += tm["trade_ms"]
            plan_ms += tm["plan_ms"]
            tl += tm["trade_launches"]
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        if os.environ.get("SDX_BENCH_DEBUG"):
            print("per-step ms:", [round(a.elapsed_time(b), 2) for a, b in ev], file=sys.stderr)
        launches = eng.total_launches() - l0
        return ms, kernel_ms, plan_ms, launches, ge, wall, tl

    sampler = ClockSampler(local)
    sampler.start()
    ms, trade_ms, plan_ms, launches, ge, wall, trade_launches = timed(step_resident, args.steps, args.warmup, 0)
    sampler.stop_flag.set()
    sampler.join()
    ms_e2e, _, _, _, ge2, _, _ = timed(step_e2e, args.steps, max(args.warmup, 1), 10_000)

    strong = strong_scaling(args, eng, torch, dist, stream, flush, modules, rows_np, w_np, S, rank, world, barrier)
    full = None if args.no_extra else full_size_configs(eng, torch, dist, rank, world, barrier)
    # back to the headline workload for the roofline peaks measured below
    eng.upload_matrix(rows_np, S)
    eng.upload_weights(w_np)

    t = torch.tensor([ms, ms_e2e, trade_ms, plan_ms], dtype=torch.float64, device="cuda")
    counts[0] = ge
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)        # the path's only exchange: exceedance counts
    ms, ms_e2e, trade_ms, plan_ms = [float(x) for x in t.tolist()]
    total_nulls = n * args.steps * world
    value = total_nulls / (ms / 1e3)
    e2e_value = total_nulls / (ms_e2e / 1e3)

    line = None
    if rank == 0:
        peaks = {}
        peaks_src = "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
            peaks_src = "measured"
        except Exception:        # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # algorithmic bytes per null: read observed + 4 row transfers per trade
        wd = wpr
        bytes_per_null = R * wd * 4 + T * 4 * wd * 4
        trade_launches = max(int(trade_launches), 1)         # k_trade launches in the timed region (sdx_timing.trade_launches)
        per_launch_nulls = n * args.steps / trade_launches   # whole waves of chains per launch
        avg_launch_ms = trade_ms / trade_launches
        popc_per_null = T * 2 * wd
        int_peaks = eng.measure_int_peaks()          # POPC / LOP3 rates measured on this box, outside the timed region
        line = {
            "metric": "curveball_nulls_scored_per_s", "value": value, "unit": "nulls/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32+f64", "data": "synthetic",
            "config": {"workload": "C1: BRAF-shape %dx%d (synthetic DepMap-20Q2 shape), fixed k=3 module, %d Curveball nulls per GPU per step, "
                                   "%d trades each, fused SuperW scoring + exceedance count" % (S, F, n, T),
                       "nulls_per_step_per_gpu": n, "trades_per_null": T, "chain_len": CHAIN_LEN, "stat": "fixed",
                       "burn_in": "chains start from a pool of %d states burnt in for %d rounds; the pool is cached per (matrix, seed): "
                                  "built once before the timed region of `value`, rebuilt every step of `e2e` (the matrix is re-uploaded)" % (POOL_SIZE, BURN_ROUNDS),
                       "pair_group": PAIR_GROUP,
                       "l2": "flushed between timed iterations (256 MiB memset); the working states of a launch (%d MB) exceed L2" % (n // 32 * 220 // 1000),
                       "sharding": "null index ranges per rank, all_reduce(SUM) of exceedance counts"},
            "e2e": {"value": e2e_value, "unit": "nulls/s",
                    "h2d_bytes_per_step": int(rows_np.nbytes + w_np.nbytes + modules.nbytes),
                    "d2h_bytes_per_step": int(8 + 8), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "issue (INT / popc pipe; not HBM, not tensor)", "kernel": "k_trade_lanes",
                         "achieved": popc_per_null * per_launch_nulls / (avg_launch_ms / 1e3), "peak": int_peaks["popc_per_s"], "unit": "popc/s",
                         "frac": popc_per_null * per_launch_nulls / (avg_launch_ms / 1e3) / int_peaks["popc_per_s"],
                         "definition": "SURVEY 8d: algorithmic popc = 2 * T * Wd = %d per null x nulls per launch / launch time (CUDA events on the "
                                       "launching stream), against the 32-bit POPC rate measured on this box by sdx_measure_int_peaks" % popc_per_null,
                         "peak_source": "measured on this box (register-only micro-kernel); nominal %.3g" % POPC_PEAK_NOMINAL,
                         "frac_of_nominal": popc_per_null * per_launch_nulls / (avg_launch_ms / 1e3) / POPC_PEAK_NOMINAL,
                         "lop3_peak_measured": int_peaks["lop3_per_s"],
                         "traffic": NCU_DRAM_BYTES_PER_NULL * per_launch_nulls,
                         "traffic_source": "ncu --set full capture of one 100000-null launch of this configuration (profiles/r02_k_trade_lanes_ncu_summary.md): "
                                           "dram__bytes_read.sum + dram__bytes_write.sum per null x nulls per launch",
                         "hbm": {"achieved_gbs": NCU_DRAM_BYTES_PER_NULL * per_launch_nulls / 1e9 / (avg_launch_ms / 1e3), "peak_gbs": hbm_peak,
                                 "frac": NCU_DRAM_BYTES_PER_NULL * per_launch_nulls / 1e9 / (avg_launch_ms / 1e3) / hbm_peak, "peak_source": peaks_src,
                                 "note": "the states of the %d groups of 32 nulls of a launch (220 KB each) do not fit L2: every row of every trade "
                                         "is staged from HBM by TMA and written back; far from the HBM roofline, which is why the bound is issue" % int(per_launch_nulls // 32)},
                         "row_bytes_per_null_algorithmic": bytes_per_null,
                         "avg_launch_ms": avg_launch_ms, "nulls_per_launch": per_launch_nulls, "trade_launches": trade_launches,
                         "kernel_share_of_step": trade_ms / ms, "plan_share_of_step": plan_ms / ms},
            "strong_scaling": strong,
            "full_size_configs": full,
            "clocks": sampler.summary(),
            "exceedances": int(counts.item()), "p_value": float(counts.item()) / total_nulls,
            "wall_s_timed_region": wall,
        }
    # ---- secondary figures + CPU baseline, rank 0 at N=1 only ------------------
    if rank == 0 and world == 1 and not args.no_extra:
        line["extra"] = secondary(eng, torch)
        cpu_v, cpu_dt, _ = cpu_nulls_per_second(dense, w, module, args.cpu_nulls, SEED)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "nulls/s", "cores": 1, "kind": "port",
                                "sample": "%d chained nulls of the same matrix/module in %.1f s, oracle/refport.py (the reference's "
                                          "Python algorithm, CPython %s)" % (args.cpu_nulls, cpu_dt, sys.version.split()[0])}
    if rank == 0:
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def strong_scaling(args, eng, torch, dist, stream, flush, modules, rows_np, w_np, S, rank, world, barrier):
    """configs[1] as BASELINE.json states it: 100 000 nulls IN TOTAL, sharded over the ranks by null index the way the
    reference shards one `-p` over processes by `-pre` (src/generate_null_matrices.py:29,82), through the product's own
    superdendrix_b200.dist.null_pvalue — the all-reduce of the counts and the read of the result are inside the timed
    region.  `resident`: matrix, weights and burn-in pool already on the device.  `e2e`: every step uploads matrix and
    weights from pinned host memory, so the pool is rebuilt — each rank burns 1/world of it, one all-gather over NVLink."""
    from superdendrix_b200 import dist as sdist
    total = NULLS_PER_STEP
    steps, warmup = max(args.steps, 3), max(args.warmup, 2)
    z_obs = eng.null_pvalue(SEED, 0, 0, modules)["z_obs"]
    out = {}

    def run(e2e, step):
        if e2e:
            eng.upload_matrix(rows_np, S)
            eng.upload_weights(w_np)
        r = sdist.null_pvalue(eng, SEED, total, modules, null0=step * total, chain_len=CHAIN_LEN, burn=BURN_ROUNDS, pool_size=POOL_SIZE,
                              pair_group=PAIR_GROUP, z_obs=None if e2e else z_obs)
        return int(r["n_ge"][0])

    for name, e2e in (("resident", False), ("e2e", True)):
        for i in range(warmup):
            run(e2e, 50_000 + i)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ge = []
        for i in range(steps):
            flush.zero_()
            barrier()
            ev[i][0].record(stream)
            ge.append(run(e2e, 60_000 + i))
            ev[i][1].record(stream)
        barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = {"ms_per_100k_nulls": float(ms.item()), "nulls_per_s": total / (float(ms.item()) / 1e3), "exceedances_first_step": ge[0]}
    out["workload"] = "C1 as stated: %d nulls in total over %d GPU(s), dist.null_pvalue (all_reduce and result read inside the timed region)" % (total, world)
    out["scaling"] = "strong"
    return out


def full_size_configs(eng, torch, dist, rank, world, barrier):
    """BASELINE configs[2..4] at FULL size through superdendrix_b200.dist on `world` GPUs (sharded by subset range / null
    range / job; collectives and result reads inside the timed region; wall time between barriers, max over ranks):
      C2  exhaustive k<=3 scan, 1 000 x ~2 000
      C3  batched screen: 500 profiles x 10 000 shared nulls of the 1 000 x ~5 000 matrix (rows = samples: k_trade, global-memory variant)
      C4  18 000 profiles: conditional permutations (3 features x 10 000 permutations each) + rank-sum"""
    from superdendrix_b200 import bitpack, dist as sdist, synth
    out = {"n_gpus": world}

    def timed(fn, reps=2):
        best, res = None, None
        for _ in range(reps):
            barrier()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best, res

    try:
        dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
        eng.upload_matrix(bitpack.pack_rows(dense.T), 1000)
        eng.upload_weights(synth.weights(1000, 1, seed=3))
        dt, r = timed(lambda: sdist.scan(eng, 0, 3, 16), reps=3)
        out["c2_scan"] = {"workload": "C2: exhaustive k<=3 scan, %d features x 1000 samples, sharded by smallest feature index" % dense.shape[1],
                          "subsets": int(r["n_scored"]), "ms": dt * 1e3, "subsets_per_s": r["n_scored"] / dt, "best": [int(x) for x in r["features"][0]],
                          "best_W": float(r["W"][0])}
    except Exception as e:      # noqa: BLE001
        out["c2_scan"] = {"error": str(e)}
    try:
        dense, _, _ = synth.feature_matrix(1000, 5000, seed=20)
        S, F = dense.shape
        n_prof, n_nulls = 500, 10000
        w = synth.weights(S, n_prof, seed=5)
        eng.upload_matrix(bitpack.pack_rows(dense.T), S)
        eng.upload_weights(w)
        rng = np.random.default_rng(2)
        order = np.argsort(-dense.sum(0), kind="stable")[:60]
        mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
        sdist.null_pvalue(eng, SEED, 256 * world, mods)                       # warm-up: pool (1472 x 640 KB) + scratch
        dt, r = timed(lambda: sdist.null_pvalue(eng, SEED, n_nulls, mods), reps=2)
        out["c3_screen"] = {"workload": "C3: %d profiles x %d shared nulls, %d x %d matrix (rows = samples, %d trades per null), pool resident" % (
                                n_prof, n_nulls, S, F, 5 * S),
                            "ms": dt * 1e3, "nulls_per_s": n_nulls / dt, "profile_null_scores_per_s": n_nulls * n_prof / dt,
                            "p_mean": float(np.mean(r["p"]))}
    except Exception as e:      # noqa: BLE001
        out["c3_screen"] = {"error": str(e)}
    try:
        dense, _, _ = workload()
        S, F = dense.shape
        n_prof = 18000
        w = synth.weights(S, n_prof, seed=11)
        eng.upload_matrix(bitpack.pack_rows(dense.T), S)
        eng.upload_weights(w)
        rng = np.random.default_rng(1)
        order = np.argsort(-dense.sum(0), kind="stable")[:40]
        mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
        pj = np.arange(n_prof, dtype=np.int32)
        sdist.condperm_batch(eng, mods[:64 * world], 1000, SEED, 0, pj[:64 * world])
        dt, r = timed(lambda: sdist.condperm_batch(eng, mods, 10000, SEED, 0, pj), reps=2)
        dt2, r2 = timed(lambda: sdist.ranksum(eng, mods, pj), reps=2)
        out["c4_genome_wide"] = {"workload": "C4: %d profiles x k=3 module: 10000 conditional permutations per feature, then rank-sum per profile" % n_prof,
                                 "condperm_ms": dt * 1e3, "permuted_scores_per_s": n_prof * 3 * 10000 / dt,
                                 "ranksum_ms": dt2 * 1e3, "ranksum_profiles_per_s": n_prof / dt2, "ranksum_p_median": float(np.median(r2["p"]))}
    except Exception as e:      # noqa: BLE001
        out["c4_genome_wide"] = {"error": str(e)}
    return out


def secondary(eng, torch):
    """Scan / max_k / cond-perm / rank-sum throughput (not the headline; device-timed by the library's own events)."""
    from superdendrix_b200 import bitpack, synth
    from superdendrix_b200.engine import SDX_STAT_MAXK
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as e:      # noqa: BLE001
            out[name] = {"error": str(e)}

    def scan_c2():
        dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
        w = synth.weights(1000, 1, seed=3)
        eng.upload_matrix(bitpack.pack_rows(dense.T), 1000)
        eng.upload_weights(w)
        eng.scan(0, 3, 8)
        ts = []
        for _ in range(3):
            r = eng.scan(0, 3, 8)
            ts.append(eng.timing()["total_ms"])
        return {"workload": "C2: exhaustive k<=3 scan, %d features x 1000 samples" % dense.shape[1], "subsets": int(r["n_scored"]),
                "ms": min(ts), "subsets_per_s": r["n_scored"] / (min(ts) / 1e3), "best_W": float(r["W"][0]),
                "best": [int(x) for x in r["features"][0]],
                "popc_roofline_frac_of_nominal": r["n_scored"] / (min(ts) / 1e3) * 32 / POPC_PEAK_NOMINAL,
                "popc_roofline_frac_of_measured": r["n_scored"] / (min(ts) / 1e3) * 32 / eng.measure_int_peaks()["popc_per_s"],
                "note": "inclusion-exclusion over a pair table does fewer word operations than the per-subset popc count the roofline "
                        "formula assumes (SURVEY 8d), so the fraction can exceed 1"}

    def maxk_c0():
        dense, w, module = workload()
        eng.upload_matrix(bitpack.pack_rows(dense.T), dense.shape[0])
        eng.upload_weights(w)
        n = 1000
        eng.null_pvalue(SEED, 0, 64, module[None, :], stat=SDX_STAT_MAXK)
        r = eng.null_pvalue(SEED, 0, n, module[None, :], stat=SDX_STAT_MAXK)
        t = eng.timing()
        F = dense.shape[1]
        per_null = F + F * (F - 1) // 2 + F * (F - 1) * (F - 2) // 6
        return {"workload": "C0: 1000 nulls, each re-optimised over all |M|<=3 (the reference's per-null ILP statistic)",
                "ms": t["total_ms"], "nulls_per_s": n / (t["total_ms"] / 1e3), "subsets_per_s": n * per_null / (t["total_ms"] / 1e3),
                "n_ge": int(r["n_ge"][0]), "z_obs": float(r["z_obs"][0])}

    def condperm_c4():
        dense, _, _ = workload()
        S, F = dense.shape
        n_prof = 2000
        w = synth.weights(S, n_prof, seed=11)
        eng.upload_matrix(bitpack.pack_rows(dense.T), S)
        eng.upload_weights(w)
        rng = np.random.default_rng(1)
        order = np.argsort(-dense.sum(0), kind="stable")[:40]
        mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
        eng.condperm_batch(mods[:64], 1000, SEED, 0, np.arange(64, dtype=np.int32))
        eng.condperm_batch(mods, 10000, SEED, 0, np.arange(n_prof, dtype=np.int32))
        ms = eng.timing()["total_ms"]
        res = eng.ranksum(mods, np.arange(n_prof, dtype=np.int32))
        rms = eng.timing()["score_ms"]
        return {"workload": "C4 slice: %d profiles x k=3 module x 10000 conditional permutations per feature; rank-sum per profile" % n_prof,
                "condperm_ms": ms, "permuted_scores_per_s": n_prof * 3 * 10000 / (ms / 1e3),
                "ranksum_ms": rms, "ranksum_profiles_per_s": n_prof / (rms / 1e3), "ranksum_p_median": float(np.median(res["p"]))}

    def nulls_c3():
        dense, _, _ = synth.feature_matrix(1000, 5000, seed=20)
        S, F = dense.shape
        n_prof = 500
        w = synth.weights(S, n_prof, seed=5)
        eng.upload_matrix(bitpack.pack_rows(dense.T), S)
        eng.upload_weights(w)
        rng = np.random.default_rng(2)
        order = np.argsort(-dense.sum(0), kind="stable")[:60]
        mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
        eng.null_pvalue(SEED, 0, 64, mods)
        n = 1024
        r = eng.null_pvalue(SEED, 0, n, mods)
        t = eng.timing()
        R, L, wpr, ras = eng.trade_shape()
        return {"workload": "C3 slice: %d x %d matrix (rows = samples, %d-bit rows, %d trades per null, matrix in L2/global), "
                            "%d nulls each scored for %d profiles" % (S, F, L, 5 * R, n, n_prof),
                "ms": t["total_ms"], "trade_ms": t["trade_ms"], "plan_ms": t["plan_ms"], "nulls_per_s": n / (t["total_ms"] / 1e3),
                "profile_null_scores_per_s": n * n_prof / (t["total_ms"] / 1e3), "n_ge_mean": float(np.mean(r["n_ge"]))}

    def files_c0():
        import shutil
        import tempfile

        from superdendrix_b200 import generate_null_matrices as gnm, i_o
        dense, samples, features = synth.feature_matrix(770, 400, seed=20)
        d = tempfile.mkdtemp(prefix="sdx_bench_")
        try:
            mut = os.path.join(d, "features.tsv")
            synth.write_features_tsv(mut, dense, samples, features)
            os.mkdir(os.path.join(d, "rand_mat"))
            n = 2000
            t0 = time.perf_counter()
            import contextlib
            with contextlib.redirect_stdout(sys.stderr):      # the CLI prints like the reference's; stdout carries the JSON line only
                gnm.main(["-m", mut, "-p", str(n), "-o", os.path.join(d, "rand_mat") + "/", "--mode", "restart", "-v", "40"])
            t_gen = time.perf_counter() - t0
            t0 = time.perf_counter()
            rows, _ = i_o.read_nulls_tsv(os.path.join(d, "rand_mat") + "/", 0, n, samples, features)
            t_read = time.perf_counter() - t0
            ok = bool((bitpack.popcount_rows(rows[0]) == dense.sum(0)).all())
        finally:
            shutil.rmtree(d, ignore_errors=True)
        return {"workload": "drop-in generator CLI: %d null TSV files of the C0 matrix written in the reference's format, then parsed back" % n,
                "generate_and_write_s": t_gen, "files_per_s": n / t_gen, "parse_s": t_read, "parsed_files_per_s": n / t_read, "margins_ok": ok}

    def cpu_secondary():
        """SURVEY 8d CPU column for the secondary metrics: the reference's Python algorithm (oracle/refport.py) and
        scipy.stats.ranksums on ONE host core, each on a bounded sample (about 2-5 s) of the same synthetic workload."""
        import random

        import scipy.stats

        from oracle import refport
        res = {"cores": 1, "kind": "port"}
        # subsets/s: SuperW over random triples of the C2 matrix
        dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
        S, F = dense.shape
        w = synth.weights(S, 1, seed=3)[0]
        profile = {i: float(w[i]) for i in range(S)}
        cases = [set(np.flatnonzero(dense[:, g]).tolist()) for g in range(F)]
        rng = np.random.default_rng(5)
        triples = rng.integers(0, F, size=(300000, 3))
        t0 = time.perf_counter()
        for a, b, c in triples:
            refport.super_w([cases[a], cases[b], cases[c]], profile)
        dt = time.perf_counter() - t0
        res["superw_subsets_per_s"] = len(triples) / dt
        res["superw_sample"] = "%d random triples of the C2 matrix in %.1f s (full scan: %.2e subsets)" % (len(triples), dt, F * (F - 1) * (F - 2) / 6)
        # permuted scores/s and rank-sum profiles/s on the C0 matrix
        dense, w1, module = workload()
        S, F = dense.shape
        samples = list(range(S))
        ev = {int(g): set(np.flatnonzero(dense[:, g]).tolist()) for g in module}
        profile = {i: float(w1[0, i]) for i in samples}
        random.seed(1)
        n_perm = 4000
        t0 = time.perf_counter()
        refport.conditional_permutation([int(g) for g in module], ev, profile, samples, n_perm)
        dt = time.perf_counter() - t0
        res["condperm_scores_per_s"] = len(module) * n_perm / dt
        res["condperm_sample"] = "%d permutations x %d features of one C0 module in %.1f s" % (n_perm, len(module), dt)
        ws = synth.weights(S, 3000, seed=11)
        union = np.zeros(S, dtype=bool)
        for g in module:
            union |= dense[:, g] != 0
        t0 = time.perf_counter()
        for p_ in range(len(ws)):
            scipy.stats.ranksums(ws[p_][union], ws[p_][~union])
        dt = time.perf_counter() - t0
        res["ranksum_profiles_per_s"] = len(ws) / dt
        res["ranksum_sample"] = "scipy.stats.ranksums on %d profiles in %.2f s" % (len(ws), dt)
        # generator rows + file writing (src/generate_null_matrices.py:75-86), 20 nulls
        import tempfile
        dense, samples_n, features = synth.feature_matrix(770, 400, seed=20)
        random.seed(2)
        lists = refport.presences(dense.astype(np.float64))
        with tempfile.TemporaryDirectory(prefix="sdx_cpu_") as d:
            t0 = time.perf_counter()
            for i in range(60):
                null = refport.curveball(dense.astype(np.float64), lists)
                rows = refport.null_rows(null, samples_n, features)
                with open(os.path.join(d, "%d.tsv" % i), "w") as fh:
                    fh.write("#Samples\tEvents\n")
                    fh.write("\n".join("\t".join(r) for r in rows))
            dt = time.perf_counter() - t0
        # the reference-faithful null statistic (per-null ILP, src/superdendrix.py:191): Gurobi is not installed, so the same model
        # solved by HiGHS (oracle/milp.py) on a few nulls is an INDICATIVE stand-in for the CPU cost of max_k
        from oracle import milp
        w0 = synth.weights(770, 1, seed=3)[0]
        random.seed(3)
        lists2 = refport.presences(dense.astype(np.float64))
        t_ilp, n_ilp = 0.0, 12
        for _ in range(n_ilp):
            null = refport.curveball(dense.astype(np.float64), lists2)
            t1 = time.perf_counter()
            milp.ilp_optimum(null, w0, 3, all_samples=False)
            t_ilp += time.perf_counter() - t1
        res["ilp_highs_nulls_per_s"] = n_ilp / t_ilp
        res["ilp_sample"] = "%d C0 nulls, k <= 3 ILP of the reference solved by HiGHS in %.1f s (Gurobi unavailable; indicative only)" % (n_ilp, t_ilp)
        res["generator_files_per_s"] = 60 / dt
        res["generator_sample"] = "60 chained nulls of the C0 matrix generated and written as TSV in %.1f s" % dt
        return res

    guarded("scan_k3_2000", scan_c2)
    guarded("cpu_secondary", cpu_secondary)
    guarded("generator_files_c0", files_c0)
    guarded("nulls_c3", nulls_c3)
    guarded("maxk_c0", maxk_c0)
    guarded("condperm_ranksum_c4", condperm_c4)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nulls", type=int, default=NULLS_PER_STEP, help="nulls per GPU per step")
    ap.add_argument("--cpu-nulls", type=int, default=300, help="size of the CPU baseline sample")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    # stdout carries the ONE JSON line and nothing else: whatever libraries print there (NCCL's version banner under NCCL_DEBUG, CLI
    # logs) goes to stderr; the line itself is written to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


_OUT = sys.stdout


def emit(line):
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


if __name__ == "__main__":
    sys.exit(main())
