#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c10
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "rc=$?"; tail -2 ${O}_smoke.txt | cut -c1-600
echo "== gemm + seq tests"; timeout 1500 python -m pytest tests/test_tma_gemm_gpu.py tests/test_seq_attention_gpu.py -q -m gpu -x 2>&1 | tail -25 | cut -c1-400
echo "== bench_seq tensor-core"; timeout 600 python scripts/bench_seq.py --steps 5 > ${O}_bench_seq_tc.json 2> ${O}_bench_seq_tc.err; echo "rc=$?"; tail -1 ${O}_bench_seq_tc.json | cut -c1-900; tail -3 ${O}_bench_seq_tc.err | cut -c1-300
echo "== bench_seq cuda-core"; BUGLAB_B200_SEQ_TC=0 timeout 600 python scripts/bench_seq.py --steps 5 > ${O}_bench_seq_simt.json 2> ${O}_bench_seq_simt.err; echo "rc=$?"; tail -1 ${O}_bench_seq_simt.json | cut -c1-900
echo "== model tests"; timeout 1500 python -m pytest tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -5 | cut -c1-400
