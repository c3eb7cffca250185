#!/bin/bash
# 2-GPU call: partition tests over NCCL / peer stores, bench --gpus 2 (strips + C4 strong scaling)
set -u
out=gpurun_out; mkdir -p $out; tag=r2i
nvidia-smi -L > $out/${tag}_gpus.txt
timeout 900 python -m pytest tests/test_partition_gpu.py -m gpu -x -q 2>&1 | tail -6 > $out/${tag}_pytest_partition.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > $out/${tag}_bench_n2.json 2> $out/${tag}_bench_n2.err
cat $out/${tag}_gpus.txt $out/${tag}_pytest_partition.log
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2i_bench_n2.json").read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","p2p_bit_identical","p2p_transport","timed_blocks_ms")})
    print("e2e", d["e2e"])
    print("c4_strong", json.dumps(d.get("c4_strong"))[:1500])
except Exception as e: print("ERR", e)
PY
tail -25 $out/${tag}_bench_n2.err
