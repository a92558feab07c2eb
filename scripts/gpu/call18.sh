#!/usr/bin/env bash
# Round-2 GPU call 18: does the side-stream overlap of the by-source edge backward cost end-to-end throughput?  Same box, alternating.
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do
for ov in 1 0; do
echo "== overlap=$ov rep $rep"; BUGLAB_B200_OVERLAP=$ov timeout 600 python bench.py --steps 20 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'ratio', round(d['e2e']['value']/d['value'],3), d['clocks']['sm_mhz'])"
done
done
