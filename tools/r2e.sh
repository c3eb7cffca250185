#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2e
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 100 > $out/${tag}_learn_200k.log 2>&1
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 10 > $out/${tag}_learn_200k_10lf.log 2>&1
timeout 200 python tools/prof_learn.py 1000000 100 > $out/${tag}_learn_1M.log 2>&1
tail -n 4 $out/${tag}_learn_200k.log; tail -n 4 $out/${tag}_learn_200k_10lf.log; tail -n 1 $out/${tag}_learn_1M.log
