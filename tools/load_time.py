#!/usr/bin/env python
"""Loader fast path: write a graph in DeepDive's binary
graph.{meta,weights,variables,factors} format and time NumbSkull.loadFGFromFile on it -- the
threaded big-endian parsers (nb_load_*: fixed-width records in parallel, graph.factors in two
passes) and the threaded compute_var_map.  The reference needs about 2 minutes for the 4096^2 Ising
graph (Python loops at numbskull.py:311-317).

    python tools/load_time.py [rows] [cols]      (CPU only; no GPU needed)"""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import numbskull_b200 as nb  # noqa: E402
from numbskull_b200 import synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cols = int(sys.argv[2]) if len(sys.argv) > 2 else rows
g = synth.ising_grid(rows, cols)
d = tempfile.mkdtemp(prefix="nb_load_")
try:
    t0 = time.perf_counter()
    synth.write_deepdive(d, *g[:4])
    t1 = time.perf_counter()
    size = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))
    ns = nb.NumbSkull(directory=d, quiet=True)
    t2 = time.perf_counter()
    ns.loadFGFromFile()
    t3 = time.perf_counter()
    fg = ns.factorGraphs[0]
    same = all(np.array_equal(a, b) for a, b in ((fg.variable["cardinality"], g[1]["cardinality"]), (fg.factor["weightId"], g[2]["weightId"]),
                                                  (fg.fmap["vid"], g[3]["vid"])))
    print(json.dumps({"graph": "ising_%dx%d" % (rows, cols), "variables": len(g[1]), "factors": len(g[2]), "file_MB": round(size / 1e6, 1),
                      "write_s": round(t1 - t0, 2), "loadFGFromFile_s": round(t3 - t2, 2), "cores": os.cpu_count(),
                      "round_trip_identical": bool(same), "reference": "about 120 s (BASELINE.md section 2)"}))
finally:
    shutil.rmtree(d, ignore_errors=True)
