#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2k
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_partition_gpu.py tests/test_host_mirrors_gpu.py -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest.log
timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 1 > $out/${tag}_learn_1M.log
for f in 1 0; do NUMBSKULL_B200_FAN_OUT=$f NB_BENCH_WORKLOADS=c4 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_fan$f.json 2> $out/${tag}_bench_fan$f.err; done
cat $out/${tag}_pytest.log $out/${tag}_learn_1M.log; cut -c1-900 $out/${tag}_c5_10M.json
python - <<'PY'
import json
for f in (1,0):
    d=json.loads(open("gpurun_out/r2k_bench_fan%d.json"%f).read().strip().splitlines()[-1])
    print("fan",f,"c2 ms", d["ms_per_step"], "c4", d["c4"].get("ms_per_step"), d["c4"].get("roofline",{}).get("frac"), d["c4"].get("timed_blocks_ms"))
PY
