#!/usr/bin/env bash
# Round-2 GPU call 4: coalesced epilogue + operand pre-scaling + side-stream edge backward; small-batch profile.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c4
echo "== tma gemm tests"; timeout 900 python -m pytest tests/test_tma_gemm_gpu.py -q > ${O}_tma.txt 2>&1; echo "rc=$?"; tail -12 ${O}_tma.txt
echo "== kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tcgen05_gpu.py -q -s > ${O}_kernels.txt 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|Error" ${O}_kernels.txt | tail -20
echo "== precision diagnosis, hidden 256"; timeout 900 python scripts/diag_precision.py 256 8 2000 > ${O}_precision.txt 2>&1; echo "rc=$?"; cat ${O}_precision.txt | cut -c1-700
echo "== micro-benchmark"; timeout 600 python scripts/bench_tma_gemm.py > ${O}_gemm_bench.jsonl 2> ${O}_gemm_bench.err; echo "rc=$?"; cat ${O}_gemm_bench.jsonl; tail -3 ${O}_gemm_bench.err
echo "== model parity tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -s > ${O}_model_tests.txt 2>&1; echo "rc=$?"; grep -E "batch 0|passed|failed|FAILED|Error|Greatest|Mismatched" ${O}_model_tests.txt | tail -30
echo "== rest of the gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_model_gpu.py --deselect tests/test_tma_gemm_gpu.py --deselect tests/test_kernels_gpu.py --deselect tests/test_tcgen05_gpu.py > ${O}_gpu_suite.txt 2>&1; echo "rc=$?"; tail -8 ${O}_gpu_suite.txt
echo "== bench (overlap on)"; timeout 1200 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json | cut -c1-2500; tail -2 ${O}_bench.err | cut -c1-300
echo "== bench (overlap off)"; BUGLAB_B200_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline > ${O}_bench_no_overlap.json 2>/dev/null; echo "rc=$?"; cut -c1-400 ${O}_bench_no_overlap.json
echo "== small-batch profile"; timeout 600 python scripts/profile_small_batch.py > ${O}_small_batch.txt 2>&1; echo "rc=$?"; tail -45 ${O}_small_batch.txt | cut -c1-220
echo "== ncu: projection kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj_kernel -s 2 -c 2 -o ${O}_proj python scripts/bench_tma_gemm.py --iters 1 --warmup 0 --shapes 256x256 --skip-old > ${O}_ncu_proj.log 2>&1; echo "rc=$?"
echo "== launch list of one bench step"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_bench.log 2>&1; echo "rc=$?"; wc -l ${O}_launches.csv
