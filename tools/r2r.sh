#!/bin/bash
out=gpurun_out; tag=r2r
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_record_parity.py tests/test_host_mirrors_gpu.py -x -q -m gpu -k "not learn" 2>&1 | tail -n 3
for r in 1 0; do NUMBSKULL_B200_REGULAR_SLICES=$r timeout 300 python bench.py --workloads c2 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('regular $r c2 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['timed_blocks_ms'])"; done
NB_NO_LEARN=1 timeout 300 python tools/bench_configs.py c4 --scale 0.25 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4 50M inf ms', d['inference_ms_per_sweep'])"
NB_NO_LEARN=1 timeout 300 python tools/bench_configs.py c5 --scale 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 10M inf ms', d['inference_ms_per_sweep'])"
