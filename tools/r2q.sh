#!/bin/bash
# hub tasks of 64 incidences, barrier-time prefetch in learning, occupancy variants of k_gibbs_tt
out=gpurun_out; tag=r2q
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "hub or learn or marginals or degenerate" 2>&1 | tail -n 3
for pf in 1 0; do NUMBSKULL_B200_LEARN_PREFETCH=$pf timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 1 | sed "s/^/prefetch $pf: /"; done
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 2 | cut -c1-330
timeout 600 python bench.py --workloads c4 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q_bench_c4.json").read().strip().splitlines()[-1])
print("full c4 ms", d["c4"].get("ms_per_step"), d["c4"].get("roofline",{}).get("frac"), d["c4"].get("timed_blocks_ms"))
PY
c4() { NB_NO_LEARN=1 timeout 300 python tools/bench_configs.py c4 --scale 0.25 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c4 50M inf ms', d['inference_ms_per_sweep'])"; }
c4 "MINB1 U4"
for v in "6 4" "8 4" "6 2" "8 2"; do set -- $v
  ( cd numbskull_b200/csrc && rm -f build/nb_sweep.o && make -j8 EXTRA="-DNB_TT_MINB=$1 -DNB_TT_UNROLL_SWEEP=$2" 2>&1 | grep -E "error" )
  c4 "MINB$1 U$2"
done
