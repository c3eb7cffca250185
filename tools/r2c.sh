#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2c
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "learn or lf or determin or cli or loadfg or reference_test or block_schedule or coin or tied" 2>&1 | tail -25 > $out/${tag}_pytest_learn.log
timeout 600 python -m pytest tests/test_record_parity.py tests/test_host_mirrors_gpu.py tests/test_partition_gpu.py -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest_rec.log
for p in 1 0; do
  NB_NO_LEARN=1 NUMBSKULL_B200_L2_PERSIST=$p timeout 200 python tools/bench_configs.py c4 --scale 0.25 > $out/${tag}_c4_50M_persist$p.json 2> $out/${tag}_c4_50M_persist$p.err
done
timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
NUMBSKULL_B200_L2_PERSIST=0 NB_BENCH_WORKLOADS=c4 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_n1_nopersist.json 2> $out/${tag}_bench_n1_nopersist.err
cat $out/${tag}_pytest_learn.log $out/${tag}_pytest_rec.log
for f in persist1 persist0; do python -c "
import json
d=json.loads(open('$out/${tag}_c4_50M_$f.json').read().strip().splitlines()[-1]); print('$f', d['inference_ms_per_sweep'], d['inference_roofline_frac'])" 2>&1 | tail -1; done
cat $out/${tag}_c5_10M.json
python - <<'PY'
import json
for f in ["r2c_bench_n1.json","r2c_bench_n1_nopersist.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, "c2 ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"].get("compact_tallies_ms_per_step"))
        if "c4" in d: print("  c4", {k:d["c4"].get(k) for k in ("ms_per_step","value","build_s")}, d["c4"].get("roofline",{}).get("frac"), d["c4"].get("unavailable"))
        if "learn" in d: print("  learn", {k:d["learn"].get(k) for k in ("ms_per_step","value","gpu_launches_per_epoch","weights_head","unavailable")}, d["learn"].get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
tail -5 $out/${tag}_bench_n1.err
