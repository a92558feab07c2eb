#!/usr/bin/env bash
# Round-2 GPU call 20: the steady-state e2e leg of bench.py (one loader pipeline through warm-up and timed steps).
set -u
for k in 6 20; do
echo "== steps $k"; timeout 600 python bench.py --steps $k --warmup 3 --skip-extras --skip-cpu-baseline 2>gpurun_out/r2c20_$k.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'ratio', round(d['e2e']['value']/d['value'],3), 'fill ms', round(d['e2e']['pipeline_fill_ms'],1), d['clocks']['sm_mhz'], d['gpu_launches'])"; tail -2 gpurun_out/r2c20_$k.err | cut -c1-200
done
