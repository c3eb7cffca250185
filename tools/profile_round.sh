#!/bin/bash
# One GPU call that produces everything profiles/ needs for a round.  Run on a B200 box:
#
#   gpurun --timeout 900 -- 'bash tools/profile_round.sh r2a [--tests]'
#
# and afterwards, in the build container:
#
#   cp gpurun_out/<tag>_* profiles/ && python tools/ncu_summary.py profiles/<tag>_gibbs_tt2_ncu_full.csv --write-traffic
#
# Every step runs under its own timeout so that a hang cannot take the box with it.
set -u
tag=${1:?usage: profile_round.sh <tag> [--tests]}
out=gpurun_out
mkdir -p $out
if [ "${2:-}" = "--tests" ]; then
    timeout 480 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_pytest_gpu.log
fi
# bench line (never taken under a profiler)
timeout 240 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
# per-call breakdown of the end-to-end step
timeout 120 python tools/e2e_profile.py > $out/${tag}_e2e_profile.txt 2>&1
# launch list of the timed region + end-to-end steps (bench.py brackets them with cudaProfilerStart/Stop)
timeout 240 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# full capture of the sweep kernel: the launches of one sweep inside the timed region
timeout 240 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_gibbs_tt2 -c 2 -f \
    -o $out/${tag}_tt2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu -i $out/${tag}_tt2.ncu-rep --page raw --csv > $out/${tag}_gibbs_tt2_ncu_full.csv 2> /dev/null
cat $out/${tag}_pytest_gpu.log 2> /dev/null
cut -c1-260 $out/${tag}_bench_n1.json
cat $out/${tag}_e2e_profile.txt
python tools/ncu_summary.py $out/${tag}_gibbs_tt2_ncu_full.csv 2>&1 | tail -3
