#!/usr/bin/env bash
# Round-2 GPU call 12: sequence models with the dense layers on the TMA GEMMs (new shapes up to 2048), tensor-core attention
# at the config-4 shape, bench_seq; kernel tests + smoke after the win-mask revert.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c12
echo "== gemm + seq tests"; timeout 1500 python -m pytest tests/test_tma_gemm_gpu.py tests/test_seq_attention_gpu.py -q -m gpu 2>&1 | tail -15 | cut -c1-400
echo "== bench_seq"; timeout 600 python scripts/bench_seq.py --steps 8 > ${O}_bench_seq.json 2> ${O}_bench_seq.err; echo "rc=$?"; tail -1 ${O}_bench_seq.json | cut -c1-900; tail -3 ${O}_bench_seq.err | cut -c1-300
echo "== bench_seq rat"; timeout 600 python scripts/bench_seq.py --steps 8 --layer-type rat > ${O}_bench_seq_rat.json 2> ${O}_bench_seq_rat.err; echo "rc=$?"; tail -1 ${O}_bench_seq_rat.json | cut -c1-900; tail -3 ${O}_bench_seq_rat.err | cut -c1-300
echo "== kernel tests"; timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "rc=$?"; tail -1 ${O}_smoke.txt | cut -c1-500
echo "== launch list of one seq step"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${O}_seq_launches.csv python scripts/bench_seq.py --steps 1 --warmup 1 > ${O}_ncu_seq.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py ${O}_seq_launches.csv > ${O}_seq_launches.md 2>/dev/null; head -32 ${O}_seq_launches.md | cut -c1-170
