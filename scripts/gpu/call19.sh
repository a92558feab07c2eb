#!/usr/bin/env bash
# Round-2 GPU call 19: final state (side-stream overlap off by default): whole -m gpu suite, smoke, the default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c19
echo "== full gpu suite"; timeout 2400 python -m pytest tests/ -q -m gpu -x -s > ${O}_tests.txt 2>&1; echo "rc=$?"; tail -2 ${O}_tests.txt | cut -c1-300
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
echo "== default bench"; SECONDS=0; timeout 2400 python bench.py > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "rc=$? after ${SECONDS}s"; python -c "
import json; d=json.loads(open('${O}_bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'steps', d['steps'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'layer', d['roofline_layer']['frac'], 'c3', d['roofline_c3']['layer']['frac'], 'shards', d['e2e_shards']['value'], 'c1', d['config1']['gpu']['value'], 'c4', d['config4_seq'].get('sequences_per_s'), 'cpu', d['cpu_baseline']['value'])"
