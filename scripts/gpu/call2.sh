#!/usr/bin/env bash
# Round-2 GPU call 2: first run of the TMA-fed tcgen05 GEMMs: unit tests (pairs, then single CTAs), micro-benchmark
# against the first-generation kernels, then the model suite and the bench on whichever configuration is healthy.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c2
echo "== tma gemm tests, CTA pairs"; timeout 900 python -m pytest tests/test_tma_gemm_gpu.py -q -k "2]" > ${O}_tma_cg2.txt 2>&1; rc2=$?; tail -25 ${O}_tma_cg2.txt
echo "== tma gemm tests, CTA pairs + weight-stationary"; timeout 900 python -m pytest tests/test_tma_gemm_gpu.py -q -k "2s]" > ${O}_tma_cg2s.txt 2>&1; rc2s=$?; tail -12 ${O}_tma_cg2s.txt
if [[ $rc2s -ne 0 ]]; then export BUGLAB_B200_TMA_BSTAT=0; echo "!! weight-stationary kernel unhealthy: streaming kernel only"; fi
echo "== tma gemm tests, single CTAs"; timeout 900 python -m pytest tests/test_tma_gemm_gpu.py -q -k "1]" > ${O}_tma_cg1.txt 2>&1; rc1=$?; tail -25 ${O}_tma_cg1.txt
if [[ $rc2 -ne 0 ]]; then
  if [[ $rc1 -eq 0 ]]; then export BUGLAB_B200_TMA_CG=1; echo "!! falling back to single CTAs"; else export BUGLAB_B200_TMA=0; echo "!! TMA GEMMs unhealthy: first-generation kernels"; fi
fi
if [[ "${BUGLAB_B200_TMA:-1}" != "0" ]]; then
  echo "== micro-benchmark (default cluster size)"; timeout 600 python scripts/bench_tma_gemm.py > ${O}_gemm_bench.jsonl 2> ${O}_gemm_bench.err; echo "rc=$?"; cat ${O}_gemm_bench.jsonl; tail -5 ${O}_gemm_bench.err
  if [[ $rc1 -eq 0 && "${BUGLAB_B200_TMA_CG:-2}" != "1" ]]; then
    echo "== micro-benchmark (single CTAs)"; BUGLAB_B200_TMA_CG=1 timeout 600 python scripts/bench_tma_gemm.py --skip-old > ${O}_gemm_bench_cg1.jsonl 2> ${O}_gemm_bench_cg1.err; echo "rc=$?"; cat ${O}_gemm_bench_cg1.jsonl
  fi
fi
echo "== model parity tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -s > ${O}_model_tests.txt 2>&1; echo "rc=$?"; grep -E "batch 0|passed|failed|FAILED|Error" ${O}_model_tests.txt | tail -30
echo "== rest of the gpu suite"; timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_model_gpu.py --deselect tests/test_tma_gemm_gpu.py > ${O}_gpu_suite.txt 2>&1; echo "rc=$?"; tail -15 ${O}_gpu_suite.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; cat ${O}_bench.json; tail -3 ${O}_bench.err
echo "== launch list of one bench step"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --profile > ${O}_ncu_bench.log 2>&1; echo "rc=$?"; wc -l ${O}_launches.csv
