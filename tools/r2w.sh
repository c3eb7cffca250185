#!/bin/bash
# 2-GPU call: partition tests (NCCL / P2P) and the N=2 bench line with the final code
out=gpurun_out; tag=r2w; N=${1:-2}
[ "${SKIP_TESTS:-0}" = 1 ] || timeout 900 python -m pytest tests/test_partition_gpu.py -x -q -m gpu 2>&1 | tail -n 4 > $out/${tag}_pytest_partition_n$N.log
cat $out/${tag}_pytest_partition_n$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err ) 2> $out/${tag}_time_n$N.txt
tail -n 3 $out/${tag}_bench_n$N.err | cut -c1-300; tail -n 3 $out/${tag}_time_n$N.txt
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "identical", d.get("p2p_bit_identical"))
    print("c4_strong", json.dumps(d.get("c4_strong"))[:700])
except Exception as e: print("no line", e)
PY
