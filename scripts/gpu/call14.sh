#!/usr/bin/env bash
# Round-2 GPU call 14: what the driver runs at round end — smoke(), the whole -m gpu suite, the reference arm, the default bench.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c14
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "rc=$?"; tail -1 ${O}_smoke.txt | cut -c1-500
echo "== full gpu suite"; timeout 2400 python -m pytest tests/ -q -m gpu -s > ${O}_tests.txt 2>&1; echo "rc=$?"; tail -3 ${O}_tests.txt | cut -c1-300
echo "== reference arm"; SECONDS=0; timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "rc=$? after ${SECONDS}s"; cut -c1-700 ${O}_bench_reference.json
echo "== default bench"; SECONDS=0; timeout 2400 python bench.py > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "rc=$? after ${SECONDS}s"; cut -c1-3000 ${O}_bench_1gpu.json; tail -3 ${O}_bench_1gpu.err | cut -c1-300
