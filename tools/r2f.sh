#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2f
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 100 2>&1 | tail -n 2 > $out/${tag}_learn_200k.log
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 10 2>&1 | tail -n 2 > $out/${tag}_learn_200k_10lf.log
timeout 200 python tools/prof_learn.py 1000000 100 2>&1 | tail -n 1 > $out/${tag}_learn_1M.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "learn or lf or determin or coin or tied or marginals or potentials" 2>&1 | tail -5 > $out/${tag}_pytest.log
NB_NO_LEARN=0 timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
cat $out/${tag}_learn_200k.log $out/${tag}_learn_200k_10lf.log $out/${tag}_learn_1M.log $out/${tag}_pytest.log $out/${tag}_c5_10M.json
