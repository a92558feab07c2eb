#!/usr/bin/env bash
# Round-2 GPU call 5: unit-table kernels (streaming), full test suite, bench, launch list, ncu of gather projection in the step.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c5
echo "== full gpu suite"; timeout 2400 python -m pytest tests -q -m gpu -x > ${O}_gpu_suite.txt 2>&1; echo "rc=$?"; tail -6 ${O}_gpu_suite.txt
echo "== micro-benchmark"; timeout 600 python scripts/bench_tma_gemm.py > ${O}_gemm_bench.jsonl 2> ${O}_gemm_bench.err; echo "rc=$?"; cat ${O}_gemm_bench.jsonl; tail -3 ${O}_gemm_bench.err
echo "== bench"; timeout 1500 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json | cut -c1-1200; tail -2 ${O}_bench.err | cut -c1-300
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 4 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "rc=$?"; cut -c1-600 ${O}_bench_reference.json
echo "== launch list of one bench step"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_bench.log 2>&1; echo "rc=$?"; wc -l ${O}_launches.csv
echo "== ncu: projection (gather) + weight gradient inside the step"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"proj_kernel|wgrad_kernel" -s 6 -c 4 -o ${O}_step_gemms python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_step.log 2>&1; echo "rc=$?"
echo "== eval parity"; timeout 1200 python scripts/eval_parity.py > ${O}_eval_parity.json 2> ${O}_eval_parity.err; echo "rc=$?"; cut -c1-1500 ${O}_eval_parity.json
