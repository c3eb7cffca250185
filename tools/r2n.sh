#!/bin/bash
# where does a C5 learning epoch go in throughput mode?
out=gpurun_out; tag=r2n
export NUMBSKULL_B200_LEARN_MODE=2
NB_NO_INF=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_cell -c 120 --csv --log-file $out/${tag}_c5_launches.csv python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2n_c5_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split('(')[0]; v=float(r[-1].replace(',','')); u=r[-2]
    agg[k][0]+=1; agg[k][1]+=v
for k,(n,t) in agg.items(): print(k,n,'launches',t/n,'avg', u)
print([ (r[4].split('(')[0][:16], r[7], r[-1]) for r in rows[:30]])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cell_thread -s 6 -c 1 -o $out/${tag}_cell_thread -f python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_full.log 2>&1
ncu -i $out/${tag}_cell_thread.ncu-rep --page details --csv > $out/${tag}_cell_thread_details.csv 2>/dev/null
ncu -i $out/${tag}_cell_thread.ncu-rep --page source --csv > $out/${tag}_cell_thread_source.csv 2>/dev/null
rm -f $out/${tag}_cell_thread.ncu-rep
( cd numbskull_b200/csrc && rm -f build/nb_learn.o && make -j8 EXTRA=-DNB_EVAL_INLINE 2>&1 | grep -E "error" )
timeout 300 python tools/bench_configs.py c5 --scale 0.2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('inline lean c5 learn ms', d['learn_ms_per_epoch'], 'launches', d['learn_launches_per_epoch'])"
