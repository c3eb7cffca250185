#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; tag=r2b
timeout 900 python -m pytest tests/test_record_parity.py tests/test_host_mirrors_gpu.py -m gpu -q 2>&1 | tail -25 > $out/${tag}_pytest_new.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_record_parity.py --deselect tests/test_host_mirrors_gpu.py 2>&1 | tail -15 > $out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
for v in 0 1 2 3 4; do
  NB_NO_LEARN=1 NUMBSKULL_B200_TT_VARIANT=$v timeout 200 python tools/bench_configs.py c4 --scale 0.25 > $out/${tag}_c4_50M_v$v.json 2> $out/${tag}_c4_50M_v$v.err
done
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gibbs_tt -c 4 -f \
    -o $out/${tag}_c4tt python tools/prof_c4.py 50000000 1 > $out/${tag}_prof_c4.log 2>&1
ncu -i $out/${tag}_c4tt.ncu-rep --page raw --csv > $out/${tag}_c4tt_raw.csv 2> /dev/null
ncu -i $out/${tag}_c4tt.ncu-rep --page source --csv > $out/${tag}_c4tt_source.csv 2> /dev/null
cat $out/${tag}_pytest_new.log $out/${tag}_pytest_gpu.log
cat $out/${tag}_bench_n1.json; tail -3 $out/${tag}_bench_n1.err
cat $out/${tag}_bench_ref.json; tail -4 $out/${tag}_bench_ref.err
for v in 0 1 2 3 4; do python -c "
import json,sys
d=json.loads(open('$out/${tag}_c4_50M_v$v.json').read().strip().splitlines()[-1]); print('variant $v', d['inference_ms_per_sweep'], d['inference_roofline_frac'], d['device_build_s'], d['host_index_s'])" 2>&1 | tail -1; done
tail -3 $out/${tag}_prof_c4.log
