#!/usr/bin/env bash
# Round-2 GPU call 11: smoke() with the audited head decisions, the per-edge win masks of the edge backward, the whole -m gpu
# suite, the 1-GPU bench (resident, then full) and the launch list of one step.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c11
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "rc=$?"; tail -2 ${O}_smoke.txt | cut -c1-700
echo "== 1-GPU bench, resident only"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline > ${O}_bench_resident.json 2>/dev/null; echo "rc=$?"; cut -c1-500 ${O}_bench_resident.json
echo "== full gpu suite"; timeout 2400 python -m pytest tests/ -q -m gpu -x -s > ${O}_tests.txt 2>&1; echo "rc=$?"; tail -4 ${O}_tests.txt | cut -c1-300
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --skip-extras --skip-cpu-baseline > ${O}_ncu_bench.log 2>&1; echo "rc=$?"
python scripts/ncu_summary.py ${O}_launches.csv > ${O}_launches.md 2>/dev/null; head -30 ${O}_launches.md | cut -c1-200
