#!/bin/bash
# C3 (labelling-function model) partitioned by candidate over N GPUs, learning + inference
out=gpurun_out; tag=r2v; N=${1:-8}; SCALE=${2:-0.5}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/bench_configs.py c3 --scale $SCALE > $out/${tag}_c3_n$N.json 2> $out/${tag}_c3_n$N.err ) 2> $out/${tag}_time.txt
tail -n 3 $out/${tag}_c3_n$N.err | cut -c1-300; tail -n 3 $out/${tag}_time.txt; free -g | head -n 2
tail -n 1 $out/${tag}_c3_n$N.json | cut -c1-900
