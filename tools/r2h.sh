#!/bin/bash
out=gpurun_out; tag=r2h
NUMBSKULL_B200_LEARN_DBG_REV=10 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_learn_cells -c 1 -f \
    -o $out/${tag}_learn python tools/prof_learn.py 100000 100 > $out/${tag}_prof_learn.log 2>&1
ncu -i $out/${tag}_learn.ncu-rep --page source --csv > $out/${tag}_learn_source.csv 2> /dev/null
tail -2 $out/${tag}_prof_learn.log
