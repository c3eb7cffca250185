#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2j
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > $out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
cat $out/${tag}_pytest_gpu.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench_n1.json").read().strip().splitlines()[-1])
print("c2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"].get("frac_dram"), "e2e", d["e2e"])
print("c4", {k:d["c4"].get(k) for k in ("ms_per_step","value","build_s","unavailable")}, d["c4"].get("roofline",{}).get("frac"))
print("learn", {k:d["learn"].get(k) for k in ("ms_per_step","value","gpu_launches_per_epoch","weights_head","unavailable")}, d["learn"].get("roofline",{}).get("frac"))
print("cpu", d.get("cpu_baseline"))
PY
tail -3 $out/${tag}_bench_n1.err; cat $out/${tag}_c5_10M.json
