"""Summarise an `ncu --page raw --csv` export of the sweep kernel(s) of one sweep.

    python tools/ncu_summary.py profiles/<tag>_ncu_full.csv [--kernel k_gibbs_tt2] [--write-traffic]

Prints a markdown table (per launch and per sweep = all captured launches of the kernel) and, with
--write-traffic, rewrites profiles/traffic.json, which bench.py reports as roofline.traffic."""
import argparse
import csv
import json
import os

UNITS = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--kernel", default="k_gibbs_tt2")
    ap.add_argument("--write-traffic", action="store_true")
    ap.add_argument("--peak-gbs", type=float, default=6535.7)
    args = ap.parse_args()
    launches = [d for d in load(args.csv) if str(d.get("Kernel Name", "")).startswith(args.kernel)]
    if not launches:
        raise SystemExit("no launch of %s in %s" % (args.kernel, args.csv))
    keys = [("gpu__time_duration.sum", "time (us)"), ("dram__bytes_read.sum", "DRAM read (B)"),
            ("dram__bytes_write.sum", "DRAM write (B)"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("smsp__inst_executed.sum", "warp instructions"),
            ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "long-scoreboard stall / issue")]
    print("| metric | " + " | ".join("launch %d" % i for i in range(len(launches))) + " |")
    print("|---|" + "---|" * len(launches))
    for k, label in keys:
        if k in launches[0]:
            print("| %s | " % label + " | ".join("%.6g" % d[k] for d in launches) + " |")
    t_us = sum(d["gpu__time_duration.sum"] for d in launches)
    traffic = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in launches)
    gbs = traffic / (t_us * 1e-6) / 1e9
    print("\nper sweep (%d launches): %.1f us, %.1f MB of DRAM traffic, %.0f GB/s = %.0f %% of %.1f GB/s"
          % (len(launches), t_us, traffic / 1e6, gbs, 100 * gbs / args.peak_gbs, args.peak_gbs))
    if args.write_traffic:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "traffic.json")
        json.dump({"source": "%s (ncu --set full, %d consecutive %s launches = one sweep)"
                   % (os.path.relpath(args.csv), len(launches), args.kernel),
                   "%s_bytes_per_sweep" % args.kernel: traffic, "kernel_time_us_per_sweep_under_ncu": t_us},
                  open(path, "w"), indent=1)
        print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
