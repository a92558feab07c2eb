#!/usr/bin/env bash
# Round-2 GPU call 21: the default bench line of the final code (what the driver runs at N=1).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c21
SECONDS=0; timeout 560 python bench.py > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err; echo "rc=$? after ${SECONDS}s"; python -c "
import json; d=json.loads(open('${O}_bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'fill', d['e2e']['pipeline_fill_ms'], 'steps', d['steps'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['ms_per_launch'], 'layer', d['roofline_layer']['frac'], 'c3', d['roofline_c3']['layer']['frac'], 'shards', d['e2e_shards']['value'], 'c1', d['config1']['gpu']['value'], 'c4', d['config4_seq'].get('sequences_per_s'), 'cpu', d['cpu_baseline']['value'])"
