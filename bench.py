#!/usr/bin/env python
"""Benchmark of the Gibbs-sweep hot path (BASELINE.json metric: factor-edge
evaluations / s and variable samples / s per sweep).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full chromatic Gibbs sweep (every colour once) over the
workload: BASELINE config 2, the 4096 x 4096 Boolean Ising grid with EQUAL
pairwise factors, inference only.  With N > 1 every rank owns a 4096-row strip
of a (4096*N) x 4096 grid (weak scaling) and exchanges its boundary rows after
every colour.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput timed
with CUDA events on the library's stream; `e2e` is the same metric through
FactorGraph.inference() with host arrays, host<->device copies inside the
timed region; `roofline` relates the algorithmic bytes of a sweep (SURVEY.md
section 8d) to the measured HBM peak; `cpu_baseline` times the CPU oracle port
of the reference algorithm on a bounded sample on this box's cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

GRID = int(os.environ.get("NB_BENCH_GRID", 4096))          # rows per GPU and columns
CPU_GRID = int(os.environ.get("NB_BENCH_CPU_GRID", 1024))   # bounded sample for the CPU arm
METRIC = "factor_edge_evals_per_sec"
UNIT = "factor-edge evals/s"


# --------------------------------------------------------------------------- helpers
def algorithmic_bytes_per_sweep(n_sampled, arities_of_edge_evals):
    """B_inf = N_v*16 + sum over edge evals (20 + 5*arity)."""
    return 16.0 * n_sampled + float((20 + 5 * arities_of_edge_evals).sum())


def ising_algorithmic_bytes(rows, cols):
    nvar = rows * cols
    edge_evals = 2 * (rows * (cols - 1) + (rows - 1) * cols)   # every factor is seen from both ends
    return 16.0 * nvar + edge_evals * (20 + 5 * 2), nvar, edge_evals


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture (or null)."""
    p = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("k_gibbs_tt2_bytes_per_sweep")
        except Exception:
            pass
    return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.stop = [], False
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4)
                          if len(r) >= 7 and r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def cpu_port_throughput(steps, warmup, grid=CPU_GRID, threads=None):
    """The CPU oracle (C port of the reference's gibbsthread sweep, Hogwild over
    contiguous variable ranges like run_pool) on a bounded grid."""
    import oracle
    from numbskull_b200 import synth
    from numbskull_b200.dataloading import assign_vtf_offsets, compute_var_map
    from numbskull_b200.numbskulltypes import VarToFactor
    threads = threads or os.cpu_count() or 1
    w, v, f, fm, dm, e = synth.ising_grid(grid, grid)
    n = assign_vtf_offsets(v)
    vm, fi = np.zeros(n, VarToFactor), np.zeros(len(fm), np.int64)
    compute_var_map(v, f, fm, vm, fi, dm)
    og = oracle.OracleGraph(w, v, f, fm, vm, fi, nthreads=threads, seed=1)
    _, nvar, edges = ising_algorithmic_bytes(grid, grid)
    og.inference(0, warmup, sample_evidence=True)
    t0 = time.perf_counter()
    og.inference(0, steps, sample_evidence=True)
    dt = time.perf_counter() - t0
    return {"value": edges * steps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%dx%d Ising grid, %d sweeps, C port of gibbsthread, %d Hogwild threads"
                      % (grid, grid, steps, threads),
            "var_samples_per_sec": nvar * steps / dt, "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    r = cpu_port_throughput(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "ising_%dx%d_equal_inference (bounded sample of the %dx%d config)"
                                   % (CPU_GRID, CPU_GRID, GRID, GRID)},
            "var_samples_per_sec": r["var_samples_per_sec"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    OUT.emit(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    from numbskull_b200 import _lib, synth
    import numbskull_b200 as nb

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # stdout carries ONE JSON line: keep NCCL's banner ("NCCL version ...", printed to stdout when
        # NCCL_DEBUG=VERSION) out of it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    L = _lib.lib()
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    rows, cols = GRID, GRID
    if world == 1:
        ns = nb.NumbSkull(quiet=True)
        ns.loadFactorGraph(*synth.ising_grid(rows, cols))
        fg = ns.factorGraphs[0]
        fg.device, fg.seed = local_rank, 20261017
        runner = None
    else:
        from numbskull_b200 import partition
        runner = partition.ising_strip_runner(rows, cols, rank, world, local_rank, seed=20261017)
        fg = runner.fg
    g = fg._device_graph()
    info = fg.device_info()
    bytes_sweep, nvar, edges = ising_algorithmic_bytes(rows, cols)
    if world == 1:
        assert edges == info["n_edges"], (edges, info["n_edges"])

    def sweeps(n):
        if runner is None:
            _lib.check(L.nb_gibbs_sweeps(g, n, 0, 1, fg.seed))
        else:
            runner.sweeps(n, burnin=False, sample_evidence=True)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        _lib.check(L.nb_synchronize(g))

    fg._upload(0, 0)
    _lib.check(L.nb_reset_counts(g))
    sweeps(warmup)
    barrier()
    l0 = C.c_int64(0)
    L.nb_launch_count(g, C.byref(l0))
    def max_over_ranks(x):
        if world == 1:
            return float(x)
        import torch.distributed as dist
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # `ncu --profile-from-start off` then lists exactly the launches of the timed region and of the
    # end-to-end steps (no-ops without a profiler)
    cudart = torch.cuda.cudart()
    with ClockSampler(local_rank) as clocks:
        barrier()
        cudart.cudaProfilerStart()
        _lib.check(L.nb_timer_start(g))
        t0 = time.perf_counter()
        sweeps(steps)
        ms = C.c_float(0)
        _lib.check(L.nb_timer_stop(g, C.byref(ms)))
        cudart.cudaProfilerStop()
        barrier()
        wall = time.perf_counter() - t0
        dev_ms = max_over_ranks(ms.value)                      # max over ranks, device clock
        l1 = C.c_int64(0)
        L.nb_launch_count(g, C.byref(l1))
        # keep the GPUs busy a little longer so that the sampler sees clocks under load
        # (same count on every rank: it is derived from the rank-agreed dev_ms)
        if dev_ms < 1500:
            sweeps(max(1, int(steps * 1500 / max(dev_ms, 1e-3)) // 4))
            barrier()

    # ---- end to end through the public API with host arrays ----
    e2e_steps = max(1, min(steps, 5))
    if runner is None:
        fg.inference(0, 1, sample_evidence=True)            # warm the transfer buffers
        barrier()
        cudart.cudaProfilerStart()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fg.inference(0, 1, sample_evidence=True)
        barrier()
        e2e_dt = (time.perf_counter() - t0) / e2e_steps
        cudart.cudaProfilerStop()
    else:
        runner.inference_e2e(1)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            runner.inference_e2e(1)
        barrier()
        e2e_dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    V, Wn = len(fg.variable), len(fg.weight)
    # bytes that cross PCIe per step: values travel as 1 byte, the tallies of a call of <= 255
    # epochs as 1 byte too (4 otherwise) + the 4-byte maximum that decides it (narrowed / widened
    # on the host next to pinned staging buffers), weights as float64
    h2d = V * 1 + Wn * 8
    d2h = V * 1 + len(fg.count) * 1 + 4

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms / steps
    total_edges, total_vars = edges * world, nvar * world
    value = total_edges * steps / (dev_ms * 1e-3)
    peak, peak_src = measured_peak()
    launches = l1.value - l0.value
    achieved = bytes_sweep * steps / (dev_ms * 1e-3) / 1e9          # per GPU
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "ising_%dx%d_equal_inference%s" % (rows * world, cols,
                                                                 "" if world == 1 else "_strips_of_%d_rows" % rows),
                   "variables": total_vars, "factor_edge_evals_per_sweep": total_edges,
                   "colors": info["n_colors"] if runner is None else runner.n_colors, "l2_policy": "inputs larger than L2 (incidence stream %.0f MB vs 126 MB)"
                   % (info["stream_words"] * 4 / 1e6),
                   "partition": "single GPU" if world == 1 else
                   ("row strips; per colour a boundary phase + NVLink halo push on a side stream, concurrent with "
                    "the interior phase" if runner.p2p and runner.split else "row strips, per-colour halo exchange")},
        "var_samples_per_sec": total_vars * steps / (dev_ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(), "peak_source": peak_src,
                     "algorithmic_bytes_per_sweep": bytes_sweep,
                     "kernel": "k_gibbs_tt2 (one launch per colour; duration = CUDA-event time of the "
                               "timed region / sweeps)"},
        "gpu_launches": int(launches if runner is None else launches),
        "clocks": clocks.summary(),
        "e2e": {"value": total_edges / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_dt,
                "api": "FactorGraph.inference(0, 1) with int64/float64 host arrays"},
        "wall_ms_timed_region": 1e3 * wall,
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_port_throughput(steps=3, warmup=1)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    OUT.emit(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


class _QuietStdout(object):
    """stdout carries exactly ONE JSON line: everything libraries print there while the benchmark
    runs (e.g. the "NCCL version ..." banner at communicator creation) is routed to stderr."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line)
        sys.stdout.flush()
        os.dup2(2, 1)


OUT = None


def main():
    global OUT
    OUT = _QuietStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
