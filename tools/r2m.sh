#!/bin/bash
# throughput-mode learning + categorical record learning: parity, then the big-cell workloads
out=gpurun_out; tag=r2m
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_record_parity.py -x -q -m gpu -k "learn or cat" 2>&1 | tail -n 5
run() { export NUMBSKULL_B200_LEARN_MODE=$1
        timeout 300 python tools/bench_configs.py c5 --scale 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode $1 c5 learn ms', d['learn_ms_per_epoch'], 'inf ms', d['inference_ms_per_sweep'], 'launches', d['learn_launches_per_epoch'])"
        timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 1
        timeout 200 python tools/bench_configs.py c4 --scale 0.05 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mode $1 c4 10M learn ms', d.get('learn_ms_per_epoch'), 'inf', d['inference_ms_per_sweep'])"
        unset NUMBSKULL_B200_LEARN_MODE; }
run 0
run 1
