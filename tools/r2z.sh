#!/bin/bash
# last call of the round: full GPU suite + the default bench line of the shipped code
out=gpurun_out; tag=r2z
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest_gpu.log
timeout 420 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
cat $out/${tag}_pytest_gpu.log; cut -c1-250 $out/${tag}_bench_n1.json; tail -n 3 $out/${tag}_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_n1.json").read().strip().splitlines()[-1])
print("c2", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "c4", d["c4"]["ms_per_step"], d["c4"]["roofline"].get("frac_dram"), "learn", d["learn"]["ms_per_step"])
print("cpu", json.dumps(d["cpu_baseline"])[:700])
print("learn cpu", json.dumps(d["learn"].get("cpu_baseline"))[:400])
PY
