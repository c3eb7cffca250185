#!/bin/bash
# round 2, first GPU call: new record-parity tests + the rest of the GPU suite, bench line, other configs
set -u
out=gpurun_out; mkdir -p $out; tag=r2a
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $out/${tag}_gpu.txt; nproc >> $out/${tag}_gpu.txt; free -g | head -2 >> $out/${tag}_gpu.txt
timeout 900 python -m pytest tests/test_record_parity.py -m gpu -x -q 2>&1 | tail -15 > $out/${tag}_pytest_records.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_record_parity.py 2>&1 | tail -8 > $out/${tag}_pytest_gpu.log
timeout 240 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
NUMBSKULL_B200_UNIFORM_SLICES=0 timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/${tag}_bench_n1_nouniform.json 2> $out/${tag}_bench_n1_nouniform.err
timeout 400 python tools/bench_configs.py c4 --scale 0.25 > $out/${tag}_c4_50M.json 2> $out/${tag}_c4_50M.err
timeout 300 python tools/bench_configs.py c5 --scale 0.2 > $out/${tag}_c5_10M.json 2> $out/${tag}_c5_10M.err
timeout 300 python tools/bench_configs.py c3 --scale 0.1 > $out/${tag}_c3_1M.json 2> $out/${tag}_c3_1M.err
cat $out/${tag}_gpu.txt $out/${tag}_pytest_records.log $out/${tag}_pytest_gpu.log
cut -c1-400 $out/${tag}_bench_n1.json; cut -c1-200 $out/${tag}_bench_n1_nouniform.json
cat $out/${tag}_c4_50M.json $out/${tag}_c5_10M.json $out/${tag}_c3_1M.json
tail -3 $out/${tag}_c4_50M.err
