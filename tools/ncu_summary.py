"""Summarise an `ncu --page raw --csv` export of the kernel(s) of one step.

    python tools/ncu_summary.py profiles/<tag>_raw.csv --kernel 'k_gibbs_tt2' [--steps 1]
                                [--traffic-key k_gibbs_tt2_bytes_per_sweep]

Prints a markdown table (per launch, and per step = all captured launches whose name matches the
regular expression, divided by --steps) and, with --traffic-key, stores the DRAM bytes per step in
profiles/traffic.json, which bench.py reports as roofline.traffic / frac_dram."""
import argparse
import csv
import json
import os
import re

UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6,
         "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def load_long(rows):
    """`ncu --csv --metrics ...` log (one row per launch and metric) -> one dict per launch."""
    hdr = rows[0]
    ik, im, iu, iv, ii = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        d = per.setdefault(r[ii], {"Kernel Name": r[ik]})
        d[r[im]] = float(r[iv].replace(",", "")) * UNITS.get(r[iu], 1.0)
    return list(per.values())


def load(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    if "Metric Name" in rows[0]:
        return load_long(rows)
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {}
        for name, unit, val in zip(hdr, units, r):
            try:
                d[name] = float(val.replace(",", "")) * UNITS.get(unit, 1.0)
            except ValueError:
                d[name] = val
        out.append(d)
    return out


KEYS = [("gpu__time_duration.sum", "time (us)"), ("dram__bytes_read.sum", "DRAM read (B)"),
        ("dram__bytes_write.sum", "DRAM write (B)"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1 sectors (global loads)"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "L1 requests (global loads)"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1TEX throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "lg-throttle stall / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier stall / issue")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--kernel", default="k_gibbs_tt2", help="regular expression on the kernel name")
    ap.add_argument("--steps", type=float, default=1.0, help="steps (sweeps / epochs) the captured launches cover")
    ap.add_argument("--traffic-key", default=None)
    ap.add_argument("--peak-gbs", type=float, default=6535.7)
    ap.add_argument("--max-columns", type=int, default=4)
    args = ap.parse_args()
    rx = re.compile(args.kernel)
    launches = [d for d in load(args.csv) if rx.search(str(d.get("Kernel Name", "")))]
    if not launches:
        raise SystemExit("no launch matching %s in %s" % (args.kernel, args.csv))
    names = sorted({str(d["Kernel Name"]).split("(")[0] for d in launches})
    print("kernels: " + ", ".join(names) + "  (%d launches, %.3g step(s))\n" % (len(launches), args.steps))
    # the longest launches as columns
    cols = sorted(launches, key=lambda d: -d["gpu__time_duration.sum"])[:args.max_columns]
    print("| metric | " + " | ".join(str(d["Kernel Name"]).split("(")[0][-24:] for d in cols) + " |")
    print("|---|" + "---|" * len(cols))
    for k, label in KEYS:
        if k in cols[0]:
            print("| %s | " % label + " | ".join("%.6g" % d[k] for d in cols) + " |")
    if "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in cols[0]:
        print("| sectors / request | " + " | ".join(
            "%.2f" % (d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"] / max(d["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"], 1))
            for d in cols) + " |")
    t_us = sum(d["gpu__time_duration.sum"] for d in launches) / args.steps
    traffic = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in launches) / args.steps
    gbs = traffic / (t_us * 1e-6) / 1e9
    print("\nper step (%d launches / %.3g steps): %.1f us under ncu, %.1f MB of DRAM traffic, %.0f GB/s = %.0f %% of %.1f GB/s"
          % (len(launches), args.steps, t_us, traffic / 1e6, gbs, 100 * gbs / args.peak_gbs, args.peak_gbs))
    if args.traffic_key:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[args.traffic_key] = traffic
        cur[args.traffic_key + "_source"] = "%s (ncu, %d launches matching /%s/ over %.3g step(s); %.1f us per step under ncu)" \
            % (os.path.relpath(args.csv), len(launches), args.kernel, args.steps, t_us)
        json.dump(cur, open(path, "w"), indent=1)
        print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
