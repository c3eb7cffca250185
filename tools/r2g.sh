#!/bin/bash
out=gpurun_out; tag=r2g
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 100 2>&1 | tail -n 3 | cut -c1-900 > $out/${tag}_learn_200k.log
NUMBSKULL_B200_LEARN_TRACE=1 timeout 200 python tools/prof_learn.py 200000 10 2>&1 | tail -n 3 | cut -c1-900 > $out/${tag}_learn_200k_10lf.log
cat $out/${tag}_learn_200k.log $out/${tag}_learn_200k_10lf.log
