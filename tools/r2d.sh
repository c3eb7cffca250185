#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2d
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tied" 2>&1 | tail -5 > $out/${tag}_pytest_tied.log
timeout 200 python tools/prof_learn.py 200000 100 > $out/${tag}_learn_200k.log 2>&1
timeout 200 python tools/prof_learn.py 1000000 100 > $out/${tag}_learn_1M.log 2>&1
timeout 200 python tools/prof_learn.py 200000 10 > $out/${tag}_learn_200k_10lf.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_learn_cells -c 1 -f \
    -o $out/${tag}_learn python tools/prof_learn.py 100000 100 > $out/${tag}_prof_learn.log 2>&1
ncu -i $out/${tag}_learn.ncu-rep --page raw --csv > $out/${tag}_learn_raw.csv 2> /dev/null
ncu -i $out/${tag}_learn.ncu-rep --page source --csv > $out/${tag}_learn_source.csv 2> /dev/null
timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
cat $out/${tag}_pytest_tied.log; tail -2 $out/${tag}_learn_200k.log $out/${tag}_learn_1M.log $out/${tag}_learn_200k_10lf.log $out/${tag}_prof_learn.log
cat $out/${tag}_c5_10M.json
