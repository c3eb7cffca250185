#!/bin/bash
out=gpurun_out; tag=r2l
run() { timeout 300 python tools/bench_configs.py c5 --scale 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c5 learn ms', d['learn_ms_per_epoch'], 'inf ms', d['inference_ms_per_sweep'])"
        timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 1
        timeout 200 python tools/bench_configs.py c4 --scale 0.05 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 c4 10M learn ms', d.get('learn_ms_per_epoch'), 'inf', d['inference_ms_per_sweep'])"; }
run noinline
( cd numbskull_b200/csrc && rm -f build/nb_learn.o build/nb_sweep.o && time make -j8 EXTRA=-DNB_EVAL_INLINE 2>&1 | grep -E "error|real" )
run inline
