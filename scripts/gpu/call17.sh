#!/usr/bin/env bash
# Round-2 GPU call 17: HEAD after the reverts — whole -m gpu suite (incl. the seq-great train/evaluate entry points), smoke, resident bench.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c17
echo "== full gpu suite"; timeout 2400 python -m pytest tests/ -q -m gpu -x > ${O}_tests.txt 2>&1; echo "rc=$?"; tail -12 ${O}_tests.txt | cut -c1-300
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "== resident bench"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'])"
