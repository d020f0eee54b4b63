#!/usr/bin/env python
"""
bench.py -- BNMF-ARD Gibbs iterations/sec on the north-star configuration
(BASELINE.json configs[2]: synthetic 20000 x 20000, 20 % observed, K = 50).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one Gibbs iteration exactly as bnmf_gibbs.py:148-172: K lambda_k
draws, K U-column updates, K V-column updates, one tau draw, trace store.
Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each number
is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

HYPER = {"alphatau": 1.0, "betatau": 1.0, "alpha0": 1.0, "beta0": 1.0}
METRIC = "BNMF-ARD Gibbs iters/sec, 20k x 20k K=50"
CACHE_NOTE = "inputs exceed L2: the packed observed entries of R read by one iteration (>= 1.9 GB) against 126 MB; no flush"


def config_of(args):
    """The same dictionary in both arms (the driver compares them)."""
    return {"workload": "BNMF Gibbs+ARD %dx%d rho=%.2f K=%d (BASELINE configs[2])" % (args.I, args.J, args.density, args.K),
            "cache": CACHE_NOTE}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--I", type=int, default=20000)
    ap.add_argument("--J", type=int, default=20000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--density", type=float, default=0.2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2000, help="side of the CPU-baseline sample block")
    return ap.parse_args()


# --------------------------------------------------------------------------- #
# synthetic data: R = U* V*^T + N(0,1), U*,V* ~ Exp(1), M ~ Bernoulli(rho)
# --------------------------------------------------------------------------- #
def synth_device(I, J, K, density, device, seed=20250):
    """Full R (float64) and M (uint8) on `device`; identical on every rank."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Us = -torch.log(1.0 - torch.rand(I, K, generator=g, device=device, dtype=torch.float64))
    Vs = -torch.log(1.0 - torch.rand(J, K, generator=g, device=device, dtype=torch.float64))
    R = Us @ Vs.T
    R += torch.randn(I, J, generator=g, device=device, dtype=torch.float64)
    M = (torch.rand(I, J, generator=g, device=device, dtype=torch.float32) < density).to(torch.uint8)
    # at least one observation per row and column (bnmf_gibbs.py:92-101 asserts it)
    ri = torch.randint(0, J, (I,), generator=g, device=device)
    M[torch.arange(I, device=device), ri] = 1
    ci = torch.randint(0, I, (J,), generator=g, device=device)
    M[ci, torch.arange(J, device=device)] = 1
    return R, M


def synth_host(I, J, K, density, seed=20250):
    rng = np.random.default_rng(seed)
    R = rng.exponential(1.0, (I, K)) @ rng.exponential(1.0, (J, K)).T + rng.standard_normal((I, J))
    M = (rng.random((I, J), dtype=np.float32) < density)
    M[np.arange(I), rng.integers(0, J, I)] = True
    M[rng.integers(0, I, J), np.arange(J)] = True
    return R, M.astype(np.float64)


# --------------------------------------------------------------------------- #
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), sampled
    through NVML (falls back to the nvidia-smi command line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:  # noqa: BLE001
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:  # noqa: BLE001
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        return [sm, mx] + ["Active" if r & bit else "Not Active"
                           for bit in (flags["hw_slowdown"], flags["hw_thermal_slowdown"], flags["sw_thermal_slowdown"],
                                       flags["sw_power_cap"])]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append([str(x) for x in self._sample_nvml()])
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def load_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get("bytes_per_launch")
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# --------------------------------------------------------------------------- #
def run_reference(args):
    """CPU arm: the reference's own bnmf_gibbs (shimmed copy under oracle/_ref) --
    or the numpy oracle port if the copy is absent -- on bounded samples of the
    workload: top-left s x s blocks of the same synthetic recipe.  Two raw points
    are measured (s = 2000 and, when the budget allows, s = 4000) and printed with
    the exponent they imply; `value` is the larger block's rate scaled by
    (s*s)/(I*J): the reference's cost per iteration is proportional to I*J
    (BASELINE.md sections 2-3)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    small = cpu_baseline(args, steps=max(1, min(args.steps, 3)), warmup=min(args.warmup, 1), budget_s=25.0)
    points = [small["point"]]
    res = small
    s_big = min(2 * small["point"]["s"], args.I, args.J)
    # one iteration on the block of twice the side costs ~4x the small block's iteration (+ its set-up)
    if s_big > small["point"]["s"] and 6.0 * small["point"]["s_per_iteration"] < 150.0 - (time.perf_counter() - t_start):
        big = cpu_baseline(args, steps=1, warmup=0, budget_s=0.0, side=s_big)
        points.append(big["point"])
        res = big
    if len(points) == 2:
        import math
        expo = math.log(points[1]["s_per_iteration"] / points[0]["s_per_iteration"]) / math.log(points[1]["s"] / float(points[0]["s"]))
        res["sample"] += "; raw points %s: time per iteration grows as s^%.2f (2.0 = proportional to I*J)" % (
            ", ".join("%dx%d: %.2f s/iteration" % (q["s"], q["s"], q["s_per_iteration"]) for q in points), expo)
        res["scaling_exponent"] = expo
    res["points"] = points
    res.pop("point", None)
    line = {"metric": METRIC, "value": res["value"], "unit": "iterations/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": config_of(args),
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def cpu_baseline(args, steps=1, warmup=0, budget_s=20.0, side=None):
    """Time the reference's CPU implementation on a bounded sample: the s x s block
    (same synthetic recipe, same K, same density), s chosen from a calibration
    iteration so that warmup+steps iterations take about `budget_s` seconds."""
    import contextlib
    import io
    try:
        from threadpoolctl import threadpool_info
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:  # noqa: BLE001
        threads = os.cpu_count() or 1
    import refimport
    ref = refimport.load()
    kind = "reference" if ref is not None else "port"

    def time_block(s, n_steps, n_warm):
        R, M = synth_host(s, s, args.K, args.density)
        np.random.seed(0)
        if ref is not None:
            m = ref.bnmf_gibbs(R, M, args.K, True, HYPER)
            with contextlib.redirect_stdout(io.StringIO()):
                m.initialise("random")
                if n_warm:
                    m.run(n_warm)
                t0 = time.perf_counter()
                m.run(n_steps)
                return time.perf_counter() - t0
        import nmf_oracle as orc
        rng = np.random.default_rng(0)
        U, V = rng.exponential(1.0, (s, args.K)), rng.exponential(1.0, (s, args.K))
        lam = np.ones(args.K)
        t0 = time.perf_counter()
        for _ in range(n_steps):
            orc.gibbs_replay_iteration(R, M, U, V, 1.0, lam, lam, np.abs(U), np.abs(V))
        return time.perf_counter() - t0

    if side is None:
        s_cal = min(400, args.I, args.J)
        per_cell = time_block(s_cal, 1, 0) / float(s_cal * s_cal)          # seconds per iteration per matrix cell
        s = int(min(args.cpu_sample, args.I, args.J, max(200, (budget_s / ((steps + warmup) * per_cell)) ** 0.5)))
    else:
        s = int(side)
    dt = time_block(s, steps, warmup)
    scale = (s * s) / float(args.I * args.J)
    return {"value": steps / dt * scale, "unit": "iterations/s", "cores": int(threads), "kind": kind,
            "point": {"s": s, "s_per_iteration": dt / steps},
            "sample": "%dx%d block of the same synthetic recipe (rho=%.2f, K=%d), %d timed iteration(s) in %.1f s; "
                      "it/s scaled by (s*s)/(I*J): the reference's cost per iteration is proportional to I*J "
                      "(BASELINE.md section 2)" % (s, s, args.density, args.K, steps, dt),
            "sample_it_per_s": steps / dt}


def parity_check(args):
    """The checker, not the thing measured: one strip of the benchmark matrix in the benchmark's packed layout (20
    entries per lane, 11 rows per set, K = 50, several persistent waves of row sets) goes through two Gibbs iterations of
    the class API with the conditional parameters of every draw kept, and the oracle -- fed the same draws -- must
    reproduce them (tests/test_gpu_configs.py::test_c3_strip_gibbs_replay).  Returns the largest relative deviations."""
    import nmf_oracle as orc
    from bnmtf_ard_b200 import bnmf_gibbs
    from bnmtf_ard_b200._lib import F_MU, F_TAU, SIDE_U, SIDE_V
    I, J, K = 400, args.J, args.K
    R, M = synth_host(I, J, K, args.density, seed=77)
    os.environ["BNMTF_TRACE_PARAMS"] = "1"
    try:
        m = bnmf_gibbs(R, M, K, True, dict(HYPER))
        m.verbose = False
        m._layout_floor = [0, 600, 0, 600]      # the full matrix's longest (row, tile) segments: same layout as the timed run
        np.random.seed(5)
        m.initialise("random")
        m.run(2)
        eng = m._engine()
        info = eng.layout_info()
        lam = m.all_lambdak[1]
        tU, mU, tV, mV, _ = orc.gibbs_replay_iteration(R, M, m.all_U[0].copy(), m.all_V[0].copy(), m.all_tau[0], lam, lam,
                                                       m.all_U[1], m.all_V[1])
        rel_tau = max(float(np.max(np.abs(eng.get_factor(sd, F_TAU) - t) / t)) for sd, t in ((SIDE_U, tU), (SIDE_V, tV)))
        rel_mu = max(float(np.max(np.abs(eng.get_factor(sd, F_MU) - mu) / np.maximum(np.abs(mu), 1.0 / np.sqrt(t))))
                     for sd, mu, t in ((SIDE_U, mU, tU), (SIDE_V, mV, tV)))
        m._engine().close()
    finally:
        del os.environ["BNMTF_TRACE_PARAMS"]
    return {"max_rel": max(rel_tau, rel_mu), "max_rel_tau": rel_tau, "max_rel_mu": rel_mu, "tolerance": 1e-9,
            "what": "Gibbs replay of a %dx%d strip, K=%d, against the oracle (conditional tau and mu of every draw of iteration 2)" % (I, J, K),
            "layout": {"v2_U": int(info["v2_U"]), "v2_npt_U": int(info["v2_npt_U"]), "v2_rows_per_set_U": int(info["v2_rows_per_set_U"])}}


# --------------------------------------------------------------------------- #
def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout carries the one JSON line only: libraries that print there (NCCL's version banner comes out on stdout at its
    # default debug level) are sent to stderr for the duration of the run
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from bnmtf_ard_b200 import _dist, _lib
    from bnmtf_ard_b200._engine import NMFEngine

    I, J, K = args.I, args.J, args.K
    R, M = synth_device(I, J, K, args.density, dev)
    r0, r1, _ = _dist.shard_bounds(I, rank, world)
    c0, c1, _ = _dist.shard_bounds(J, rank, world)
    Rrow, Mrow = R[r0:r1].contiguous(), M[r0:r1].contiguous()
    Rcol, Mcol = R[:, c0:c1].contiguous(), M[:, c0:c1].contiguous()
    torch.cuda.synchronize()
    # the same run loop at every N: eager launches (no CUDA graph: multi-GPU runs cannot capture their NCCL calls; at 14 ms
    # per iteration the launch overhead is not measurable) and no per-sweep event synchronisation inside the timed region
    os.environ["BNMTF_TIME_SWEEPS"] = "0"
    os.environ["BNMTF_NO_GRAPH"] = "1"
    eng = NMFEngine.from_device_blocks(I, J, K, Rrow.data_ptr(), Mrow.data_ptr(), Rcol.data_ptr(), Mcol.data_ptr(),
                                       device=local)
    del Rrow, Mrow, Rcol, Mcol
    if rank != 0 or args.no_e2e:
        del R, M
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # state: lambda_k = alpha0/beta0, U,V ~ Exp(lambda), tau = 1  (initialise('random'))
    eng.set_hyper(True, **HYPER)
    eng.set_scalars(1.0, np.ones(K))
    eng.init_factors(_lib.GIBBS, True, 12345)

    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        parity = parity_check(args)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.run(_lib.GIBBS, args.warmup, seed=1, traces=False)
    launches0 = eng.launch_count()
    barrier()
    sampler.rows = []
    t0 = time.perf_counter()
    res = eng.run(_lib.GIBBS, args.steps, seed=2, traces=False)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.summary()
    dev_s = float(res["seconds"][-1])
    t = torch.tensor([dev_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s = float(t.item())
    launches = eng.launch_count() - launches0
    perf = eng.performances(res["stats"][-1])
    info = eng.layout_info()
    # roofline leg, outside the timed region: a few more iterations with every sweep launch bracketed by CUDA events
    eng.time_sweeps(True)
    eng.run(_lib.GIBBS, max(2, min(5, args.steps)), seed=4, traces=False)
    sweeps = eng.sweep_times()
    eng.time_sweeps(False)
    barrier()

    # ---- end to end through the reference-facing call with HOST buffers: the state
    # (U, V, tau, lambda_k) is uploaded from numpy, run() executes on the device and
    # every iteration's draws (all_U[it], all_V[it]) and statistics are copied back
    # to host numpy arrays -- exactly what bnmf_gibbs.run() of this package does.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(3, min(args.steps, 100))
        U_h = eng.get_factor(_lib.SIDE_U, _lib.F_VALUE)
        V_h = eng.get_factor(_lib.SIDE_V, _lib.F_VALUE)
        barrier()
        t0 = time.perf_counter()
        eng.set_factor(_lib.SIDE_U, _lib.F_VALUE, U_h)
        eng.set_factor(_lib.SIDE_V, _lib.F_VALUE, V_h)
        eng.set_scalars(1.0, np.ones(K))
        t_up = time.perf_counter() - t0
        # every rank holds the same chain: the draws are delivered to host memory once, by rank 0 (the other
        # ranks read back their per-iteration statistics only)
        r2 = eng.run(_lib.GIBBS, e2e_steps, seed=3, traces=(rank == 0))
        _ = (float(r2["all_U"][-1, 0, 0]) if rank == 0 else 0.0) + float(r2["stats"][-1, 0])
        t_run = time.perf_counter() - t0
        barrier()
        e2e_s = time.perf_counter() - t0
        if os.environ.get("BNMTF_RUN_TIMING"):
            print("[e2e-probe rank %d] state upload %.1f ms, run() returned at %.1f ms, barrier passed at %.1f ms"
                  % (rank, 1e3 * t_up, 1e3 * t_run, 1e3 * e2e_s), file=sys.stderr)
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
        e2e = {"value": e2e_steps / e2e_s, "unit": "iterations/s", "steps": e2e_steps,
               "h2d_bytes_per_step": int(((I + J) * K * 8 + (K + 1) * 8) / e2e_steps),
               "d2h_bytes_per_step": int((I + J) * K * 8 + K * 8 + 16 * 8),
               "note": "engine.run() through the C ABI with host numpy state and host trace buffers (pageable)" +
                       ("" if world == 1 else "; the draws go to host memory on rank 0 only (all ranks hold the same chain; "
                        "with every rank storing its own copy the 8-GPU rate is host-memory bound: profiles/r01_final_bench_n8.json)")}
        # same call with the posterior means accumulated on the device (summarise_on_device):
        # no all_U / all_V; the means of the second half of the chain are read back at the end
        eng.set_summary(e2e_steps // 2, 1)
        barrier()
        t0 = time.perf_counter()
        eng.set_factor(_lib.SIDE_U, _lib.F_VALUE, U_h)
        eng.set_factor(_lib.SIDE_V, _lib.F_VALUE, V_h)
        eng.set_scalars(1.0, np.ones(K))
        r3 = eng.run(_lib.GIBBS, e2e_steps, seed=3, traces=False)
        mU, mV = eng.get_summary(_lib.SUM_FACTOR_U), eng.get_summary(_lib.SUM_FACTOR_V)
        _ = float(mU[0, 0]) + float(mV[0, 0]) + float(r3["stats"][-1, 0])
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        eng.set_summary(0, 0)
        e2e["device_summary"] = {"value": e2e_steps / float(t.item()), "unit": "iterations/s",
                                 "d2h_bytes_per_step": int(((I + J) * K * 8) / e2e_steps + K * 8 + 16 * 8),
                                 "note": "same call, draws averaged on the device instead of copied out"}
        del mU, mV, r3
    sampler.stop_flag = True

    if rank == 0:
        peak, peak_src = load_peaks()
        value = args.steps / dev_s
        line = {"metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_of(args),
                "details": {"layout": info, "wall_s": wall, "train_mse_last": perf["MSE"],
                            "packed_gb_per_iteration": ((info["v2_slots_U"] + info["v2_slots_V"]) * 10 if info.get("v2_U") else
                                                        (info["slots_U"] + info["slots_V"]) * 12) / 1e9,
                            "value_covers": "lambda_k draws, U sweep, V sweep, tau draw and statistics of every iteration, on the "
                                            "device (multi-GPU: plus the all-gathers); the per-iteration copy of the draws into the "
                                            "trace buffers (a 16 MB device-to-device copy, < 0.1 % of an iteration) and their "
                                            "delivery to host memory are part of e2e, not of value",
                            "run_loop": "eager launches at every N (BNMTF_NO_GRAPH=1), no per-sweep synchronisation"},
                "gpu_launches": launches, "clocks": clocks}
        # roofline of the dominant kernel (row sweep): algorithmic DRAM bytes of OUR formulation
        if sweeps["n_U"] > 0:
            ms = (sweeps["ms_U"] + sweeps["ms_V"]) / (sweeps["n_U"] + sweeps["n_V"])
            if info.get("v2_U") and info.get("v2_V"):   # cluster-tile layout: f64 value + u16 column per slot, read once
                bytes_per_launch = 0.5 * (info["v2_slots_U"] + info["v2_slots_V"]) * 10.0 + (I + J) * K * 8 * 1.5
            else:
                bytes_per_launch = 0.5 * (info["slots_U"] + info["slots_V"]) * 12.0 + (I + J) * K * 8 * 1.5
            ach = bytes_per_launch / (ms * 1e-3) / 1e9
            northstar_bytes = 2 * K * 16.125 * I * J
            line["roofline"] = {"measured_in": "a separate pass of %d iterations after the timed region, each sweep launch bracketed "
                                               "by CUDA events on the engine's stream" % sweeps["n_U"],
                                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                "traffic": load_traffic(), "peak_source": peak_src,
                                "kernel": {0: "sweep_kernel", 1: "sweep2_kernel", 4: "sweep4_kernel", 5: "sweep5_kernel"}[int(info.get("v2_U", 0))],
                                "kernel_ms": ms, "kernel_share_of_step": 2.0 * ms / (1e3 * dev_s / args.steps),
                                "north_star_equiv_gbs": northstar_bytes * value / 1e9,
                                "north_star_equiv_frac": northstar_bytes * value / 1e9 / peak}
            if info.get("v2_U") and info.get("v2_V"):
                # what the cluster sweeps actually lean on: the shared-memory pipe.  Algorithmic 8-byte gathers per launch
                # over the OBSERVED entries (residual pass K + update passes 2K per entry; padded slots are not useful work)
                # at 128 B per wavefront, against one wavefront per SM and clock at the SM clock sampled during the run.
                nnz = 0.5 * (info["nnz_U"] + info["nnz_V"])
                wf = nnz * 3 * K * 8.0 / 128.0
                sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
                wf_peak = sm_count * float(clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) * 1e6
                line["roofline"]["onchip"] = {"bound": "shared-memory wavefronts", "achieved": wf / (ms * 1e-3) / 1e9,
                                              "peak": wf_peak / 1e9, "unit": "Gwavefronts/s",
                                              "frac": wf / (ms * 1e-3) / wf_peak,
                                              "padding": 0.5 * (info["v2_slots_U"] + info["v2_slots_V"]) / nnz,
                                              "note": "algorithmic gathers of the observed entries only (nnz x 3K x 8 B)"}
        line["e2e"] = e2e
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
