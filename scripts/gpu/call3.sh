#!/usr/bin/env bash
# Round-2 GPU call 3: weight-stationary kernel, precision diagnosis at H=256, ncu captures of the new kernels, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c3
echo "== tma gemm tests (pairs, pairs + stationary)"; timeout 900 python -m pytest tests/test_tma_gemm_gpu.py -q -k "2] or 2s]" > ${O}_tma.txt 2>&1; rc=$?; tail -15 ${O}_tma.txt
if [[ $rc -ne 0 ]]; then export BUGLAB_B200_TMA_BSTAT=0; echo "!! stationary kernel switched off for the rest of this call"; fi
echo "== kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -s > ${O}_kernels.txt 2>&1; echo "rc=$?"; grep -E "winners differing|passed|failed|FAILED|Error" ${O}_kernels.txt | tail -30
echo "== precision diagnosis, hidden 256"; timeout 900 python scripts/diag_precision.py 256 8 2000 > ${O}_precision.txt 2>&1; echo "rc=$?"; cat ${O}_precision.txt | cut -c1-900
echo "== micro-benchmark"; timeout 600 python scripts/bench_tma_gemm.py > ${O}_gemm_bench.jsonl 2> ${O}_gemm_bench.err; echo "rc=$?"; cat ${O}_gemm_bench.jsonl; tail -3 ${O}_gemm_bench.err
echo "== model parity tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -s > ${O}_model_tests.txt 2>&1; echo "rc=$?"; grep -E "batch 0|passed|failed|FAILED|Error|Greatest|Mismatched" ${O}_model_tests.txt | tail -30
echo "== ncu: projection kernels"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj -s 2 -c 3 -o ${O}_proj python scripts/bench_tma_gemm.py --iters 1 --warmup 0 --shapes 256x256 --skip-old > ${O}_ncu_proj.log 2>&1; echo "rc=$?"
echo "== ncu: weight gradient"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 1 -c 1 -o ${O}_wgrad python scripts/bench_tma_gemm.py --iters 1 --warmup 0 --shapes 256x256 --skip-old > ${O}_ncu_wgrad.log 2>&1; echo "rc=$?"
echo "== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json | cut -c1-3000; tail -3 ${O}_bench.err | cut -c1-400
echo "== ncu: edge backward kernels inside one step"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:edge_bwd -s 4 -c 2 -o ${O}_edge_bwd python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_edge_bwd.log 2>&1; echo "rc=$?"
echo "== launch list of one bench step"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_bench.log 2>&1; echo "rc=$?"; wc -l ${O}_launches.csv
ls -la gpurun_out/*.ncu-rep 2>/dev/null
