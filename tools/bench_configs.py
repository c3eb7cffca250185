#!/usr/bin/env python
"""Throughput of the BASELINE.json configs other than the headline one
(bench.py covers config 2).  Prints one JSON line per config:

    python tools/bench_configs.py c3 [--scale 0.1]    # LF model, learning + marginals
    python tools/bench_configs.py c4 [--scale 0.1]    # KBC-style Boolean graph, inference
    python tools/bench_configs.py c5 [--scale 0.1]    # categorical card 16, learning + inference

`--scale` multiplies the number of variables of the named config (1.0 = the
BASELINE size).  Algorithmic bytes follow SURVEY.md section 8(d)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numbskull_b200 as nb  # noqa: E402
from numbskull_b200 import _lib, synth  # noqa: E402

PEAK = 6535.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def device_time(fg, fn):
    L, g = _lib.lib(), fg._g
    _lib.check(L.nb_synchronize(g))
    _lib.check(L.nb_timer_start(g))
    fn()
    ms = C.c_float(0)
    _lib.check(L.nb_timer_stop(g, C.byref(ms)))
    return ms.value * 1e-3


def algorithmic_bytes(fg, sampled_mask, learn=False):
    """B_inf = 16 N_v + sum_{edge evals} (20 + 5 arity) (+1 per member of categorical factors)."""
    var, fac, vmap, fi = fg.variable, fg.factor, fg.vmap, fg.factor_index
    nb_ = np.where(var["dataType"] == 0, 1, var["cardinality"]).astype(np.int64)
    owner = np.repeat(np.arange(len(var)), nb_)                       # variable of every bucket
    keep = sampled_mask[owner]
    lens = vmap["factor_index_length"].astype(np.int64)
    offs = vmap["factor_index_offset"].astype(np.int64)
    total_edges = int(lens[keep].sum())
    # arity of every bucket entry of the sampled variables
    idx = np.repeat(offs[keep], lens[keep]) + (np.arange(total_edges) - np.repeat(np.cumsum(lens[keep]) - lens[keep], lens[keep]))
    ar = fac["arity"][fi[idx]].astype(np.int64)
    cat = np.isin(fac["factorFunction"][fi[idx]], (12, 14, 15, 16, 17))
    b = 16.0 * int(sampled_mask.sum()) + float((20 + 5 * ar + cat * ar).sum())
    return b, total_edges


def report(name, fg, info, extra):
    line = {"config": name, "variables": int(len(fg.variable)), "factors": int(len(fg.factor)),
            "edges": int(info["n_edges"]), "colors": int(info["n_colors"]), "wide_headers": int(info["wide_headers"]),
            "thread_rows": int(info["n_thread_rows"]), "warp_rows": int(info["n_warp_rows"]),
            "device_GB": round(info["device_bytes"] / 1e9, 2), "jp_rounds": int(info["jp_rounds"]), "peak_GBs": PEAK}
    line.update(extra)
    print(json.dumps(line))
    sys.stdout.flush()


def load(g):
    t0 = time.perf_counter()
    ns = nb.NumbSkull(quiet=True)
    ns.loadFactorGraph(*g)
    fg = ns.factorGraphs[0]
    fg.seed = 12345
    t1 = time.perf_counter()
    fg._device_graph()
    t2 = time.perf_counter()
    bad = C.c_int64(-1)
    _lib.check(_lib.lib().nb_graph_check_coloring(fg._g, C.byref(bad)))
    return fg, {"host_index_s": round(t1 - t0, 2), "device_build_s": round(t2 - t1, 2), "color_conflicts": int(bad.value)}


def inference_rate(fg, sweeps, sample_evidence=True):
    L, g = _lib.lib(), fg._g
    fg._sync_device(0, 0)
    _lib.check(L.nb_reset_counts(g))
    _lib.check(L.nb_gibbs_sweeps(g, 3, 1, int(sample_evidence), fg.seed))
    return device_time(fg, lambda: _lib.check(L.nb_gibbs_sweeps(g, sweeps, 0, int(sample_evidence), fg.seed))) / sweeps


def learn_rate(fg, epochs, stepsize, reg, reg_param, lne):
    L, g = _lib.lib(), fg._g
    fg._sync_device(0, 0)

    def run(n):
        s = C.c_double(stepsize)
        _lib.check(L.nb_learn_sweeps(g, n, C.byref(s), 1.0, reg, reg_param, 1.0, int(lne), fg.seed, 0))
    run(1)
    l0, l1 = C.c_int64(0), C.c_int64(0)
    L.nb_launch_count(g, C.byref(l0))
    dt = device_time(fg, lambda: run(epochs)) / epochs
    L.nb_launch_count(g, C.byref(l1))
    return dt, (l1.value - l0.value) / epochs


def c3(scale):
    copies, n_lf = max(1000, int(10_000_000 * scale)), 100
    rng = np.random.default_rng(1003)
    g = synth.lf_model(copies, n_lf, rng)
    acc_true = None
    fg, times = load(g)
    info = fg.device_info()
    every = fg.variable["isEvidence"] != 4
    b_free, edges = algorithmic_bytes(fg, every)
    b_evid, _ = algorithmic_bytes(fg, fg.variable["isEvidence"] != 1)
    dt, launches = learn_rate(fg, 3, 1e-4, 1, 0.01, True)
    fg._stale.add('weight_value')
    b_learn = b_free + (b_evid - 16.0 * int((fg.variable["isEvidence"] != 1).sum()) * 0.5) + 12.0 * edges
    dti = inference_rate(fg, 10, sample_evidence=False)
    q = fg.variable["isEvidence"] == 0
    b_inf, e_inf = algorithmic_bytes(fg, q)
    report("c3_lf_%dx%d" % (copies, n_lf), fg, info, dict(
        times, learn_ms_per_epoch=1e3 * dt, learn_edge_evals_per_s=edges / dt, learn_launches_per_epoch=launches,
        learn_roofline_frac=b_learn / dt / 1e9 / PEAK,
        inference_ms_per_sweep=1e3 * dti, inference_edge_evals_per_s=e_inf / dti,
        inference_roofline_frac=b_inf / dti / 1e9 / PEAK,
        weights_head=[round(float(x), 4) for x in fg.weight_value[0][:6]]))


def c4(scale):
    nvar = max(10000, int(200_000_000 * scale))
    g = synth.kbc_fast(nvar, seed=1004)
    fg, times = load(g)
    info = fg.device_info()
    every = fg.variable["isEvidence"] != 4
    # every variable is sampled and no factor repeats a member: each factor is seen once from each
    # member, so B_inf = 16 N_v + sum_f arity_f (20 + 5 arity_f) without materialising the edge list
    ar = fg.factor["arity"].astype(np.int64)
    b, edges = 16.0 * int(every.sum()) + float((ar * (20 + 5 * ar)).sum()), int(info["n_edges"])
    del ar
    dt = inference_rate(fg, 10)
    extra = dict(times, inference_ms_per_sweep=1e3 * dt, inference_edge_evals_per_s=edges / dt,
                 var_samples_per_s=int(every.sum()) / dt, inference_roofline_frac=b / dt / 1e9 / PEAK,
                 algorithmic_GB_per_sweep=b / 1e9)
    if not os.environ.get("NB_NO_LEARN"):
        dtl, launches = learn_rate(fg, 1, 0.01, 2, 0.01, False)
        extra.update(learn_ms_per_epoch=1e3 * dtl, learn_edge_evals_per_s=edges / dtl, learn_launches_per_epoch=launches)
    report("c4_kbc_%d" % nvar, fg, info, extra)


def c4_partitioned(scale):
    """C4 across the ranks of a torchrun launch: every rank builds the same global graph
    (same seed), keeps its block (owner-computes) and exchanges boundary values per colour."""
    import torch
    import torch.distributed as dist
    from numbskull_b200 import partition
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
    nvar = max(10000, int(200_000_000 * scale))
    w, v, f, fm, dm, e = synth.kbc(nvar, np.random.default_rng(1004))
    t0 = time.perf_counter()
    run = partition.partition_graph(w, v, f, fm, rank, world, local, seed=12345)
    del w, v, f, fm
    build_s = time.perf_counter() - t0
    fg = run.fg
    info = fg.device_info()
    L, g = _lib.lib(), fg._g
    fg._upload(0, 0, evid=False)
    run.sweeps(3, True, True)
    torch.cuda.synchronize()
    dist.barrier()
    sweeps = 10
    _lib.check(L.nb_timer_start(g))
    run.sweeps(sweeps, False, True)
    ms = C.c_float(0)
    _lib.check(L.nb_timer_stop(g, C.byref(ms)))
    t = torch.tensor([ms.value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_loc = torch.tensor([float(fg.color_edges().sum()), float(run.n_owned), float(len(fg.variable) - run.n_owned),
                          float(run.halo_bytes_per_sweep)], device="cuda", dtype=torch.float64)
    dist.all_reduce(e_loc)
    if rank == 0:
        dt = float(t.item()) * 1e-3 / sweeps
        print(json.dumps({"config": "c4_kbc_%d_partitioned" % nvar, "n_gpus": world, "variables": nvar,
                          "edges": int(e_loc[0].item()), "ghost_variables_total": int(e_loc[2].item()),
                          "halo_values_per_sweep_total": int(e_loc[3].item()), "colors": run.n_colors,
                          "jp_rounds": run.jp_rounds, "p2p_halo": bool(run.p2p), "build_s": round(build_s, 1),
                          "device_GB_rank0": round(info["device_bytes"] / 1e9, 2),
                          "inference_ms_per_sweep": 1e3 * dt, "inference_edge_evals_per_s": e_loc[0].item() / dt,
                          "var_samples_per_s": e_loc[1].item() / dt}))
    dist.barrier()
    dist.destroy_process_group()


def c3_partitioned(scale):
    """C3 across the ranks of a torchrun launch: every rank builds its own candidates (no ghosts);
    learning sums the ranks' weight deltas once per epoch (device-side all-reduce), then marginals."""
    import torch
    import torch.distributed as dist
    from numbskull_b200 import partition
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=900))
    copies, n_lf = max(1000, int(10_000_000 * scale)), 100
    t0 = time.perf_counter()
    loc, n_global = partition.lf_block(copies, n_lf, rank, world)
    run = partition.PartitionedGibbs(loc, n_global, rank, world, local, seed=12345)
    build_s = time.perf_counter() - t0
    fg = run.fg
    L, g = _lib.lib(), fg._g
    # weights tied across all ranks: the ranks' deltas are averaged (see PartitionedGibbs.learn)
    sync = os.environ.get("NB_WEIGHT_SYNC", "mean")
    run.learn(0, 1, 1e-4, 1.0, 1, 0.01, 1.0, learn_non_evidence=True, weight_sync=sync)            # warm-up epoch
    torch.cuda.synchronize()
    dist.barrier()
    epochs = 3
    t1 = time.perf_counter()
    run.learn(0, epochs, 1e-4, 1.0, 1, 0.01, 1.0, learn_non_evidence=True, weight_sync=sync)
    torch.cuda.synchronize()
    dist.barrier()
    dt = (time.perf_counter() - t1) / epochs
    t2 = time.perf_counter()
    sweeps = 10
    marg = run.inference(2, sweeps, sample_evidence=False)
    torch.cuda.synchronize()
    dist.barrier()
    dti = (time.perf_counter() - t2) / (sweeps + 2)
    per = 1 + n_lf
    # marginals come in the reference's cstart layout (1 entry per Boolean y, 3 per cardinality-3 LF variable)
    y = np.asarray(marg)[np.asarray(fg.cstart[0:run.n_owned:per], np.int64)] if len(marg) else np.zeros(0)
    stats = torch.tensor([float(y.sum()), float(len(y))], device="cuda", dtype=torch.float64)
    dist.all_reduce(stats)
    if rank == 0:
        edges = copies * (1 + 2 * n_lf)
        w = fg.weight_value[0]
        print(json.dumps({"config": "c3_lf_%dx%d_partitioned" % (copies, n_lf), "n_gpus": world, "variables": n_global,
                          "edges": edges, "colors": run.n_colors, "build_s": round(build_s, 1), "weight_sync": sync,
                          "device_GB_rank0": round(fg.device_info()["device_bytes"] / 1e9, 2),
                          "learn_ms_per_epoch": 1e3 * dt, "learn_edge_evals_per_s": edges / dt,
                          "inference_ms_per_sweep_wall": 1e3 * dti, "mean_marginal_y": stats[0].item() / max(stats[1].item(), 1),
                          "weights_head": [round(float(x), 4) for x in w[:6]], "weights_finite": bool(np.isfinite(w).all())}))
    dist.barrier()
    dist.destroy_process_group()


def c5(scale):
    nvar = max(10000, int(50_000_000 * scale))
    g = synth.categorical(nvar, 16, 3, np.random.default_rng(1005))
    fg, times = load(g)
    info = fg.device_info()
    every = fg.variable["isEvidence"] != 4
    b, edges = algorithmic_bytes(fg, every)
    dt = inference_rate(fg, 10)
    dtl, launches = learn_rate(fg, 2, 0.01, 2, 0.01, False)
    report("c5_cat16_%d" % nvar, fg, info, dict(
        times, inference_ms_per_sweep=1e3 * dt, inference_edge_evals_per_s=edges / dt,
        var_samples_per_s=int(every.sum()) / dt, inference_roofline_frac=b / dt / 1e9 / PEAK,
        learn_ms_per_epoch=1e3 * dtl, learn_edge_evals_per_s=edges / dtl, learn_launches_per_epoch=launches))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=0.1)
    a = ap.parse_args()
    if a.config == "c3" and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        c3_partitioned(a.scale)
    elif a.config == "c4" and int(os.environ.get("WORLD_SIZE", 1)) > 1:
        c4_partitioned(a.scale)
    else:
        {"c3": c3, "c4": c4, "c5": c5}[a.config](a.scale)
