#!/bin/bash
# bit-packed value mirror: identity test, then C4 at 50 M and at the full size, with and without
out=gpurun_out; tag=r2s
timeout 600 python -m pytest tests/test_record_parity.py tests/test_gpu_parity.py -x -q -m gpu -k "mirror or hub or marginals_match or ising_full" 2>&1 | tail -n 3
for m in -1 0; do NUMBSKULL_B200_BIT_MIRROR=$m NB_NO_LEARN=1 timeout 300 python tools/bench_configs.py c4 --scale 0.25 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mirror $m: c4 50M inf ms', d['inference_ms_per_sweep'])"; done
for m in -1 0; do NUMBSKULL_B200_BIT_MIRROR=$m timeout 600 python bench.py --workloads c4 --no-cpu-baseline 2>/dev/null > $out/${tag}_bench_c4_mirror$m.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_c4_mirror$m.json").read().strip().splitlines()[-1])
print("mirror $m: full c4 ms", d["c4"].get("ms_per_step"), d["c4"].get("roofline",{}).get("frac"), d["c4"].get("timed_blocks_ms"), d["c4"].get("mean_marginal"))
PY
done
