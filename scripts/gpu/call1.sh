#!/usr/bin/env bash
# Round-2 GPU call 1: everything round 1 prepared but never ran on a B200, plus the new whole-model parity tests.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
echo "== seq attention GPU parity" ; BUGLAB_B200_SEQ_GPU=1 timeout 600 python -m pytest tests/test_seq_attention_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/r2c1_seq_tests.txt
echo "== model parity tests" ; timeout 1200 python -m pytest tests/test_model_gpu.py -q -s 2>&1 | tail -40 | tee gpurun_out/r2c1_model_tests.txt
echo "== rest of the gpu suite" ; timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_model_gpu.py 2>&1 | tail -8 | tee gpurun_out/r2c1_gpu_suite.txt
echo "== eval parity" ; timeout 900 python scripts/eval_parity.py > gpurun_out/r2c1_eval_parity.json 2> gpurun_out/r2c1_eval_parity.err ; echo "rc=$?" ; cat gpurun_out/r2c1_eval_parity.json
echo "== e2e shards" ; timeout 600 python scripts/bench_e2e_shards.py > gpurun_out/r2c1_e2e_shards.json 2> gpurun_out/r2c1_e2e_shards.err ; echo "rc=$?" ; cat gpurun_out/r2c1_e2e_shards.json
echo "== seq bench" ; timeout 600 python scripts/bench_seq.py > gpurun_out/r2c1_bench_seq.json 2> gpurun_out/r2c1_bench_seq.err ; echo "rc=$?" ; cat gpurun_out/r2c1_bench_seq.json
echo "== bench" ; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err ; echo "rc=$?" ; cat gpurun_out/r2c1_bench.json
