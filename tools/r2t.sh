#!/bin/bash
# N-GPU validation of bench.py (weak-scaling strips, bit-identity assert, C4 strong scaling)
out=gpurun_out; tag=r2t; N=${1:-8}
nvidia-smi -L | head -n 8 > $out/${tag}_gpus.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err ) 2> $out/${tag}_time_n$N.txt
tail -n 5 $out/${tag}_bench_n$N.err | cut -c1-300
cat $out/${tag}_time_n$N.txt | tail -n 3
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "identical", d.get("p2p_bit_identical"))
    print("c4_strong", json.dumps(d.get("c4_strong"))[:900])
except Exception as e: print("no line", e)
PY
