#!/usr/bin/env bash
# Round-2 GPU call 15: the row sum on the side stream under the S-table weight gradient — parity tests, then the resident bench
# with and without the side-stream overlaps on the same box; seq tests + bench with the 16-byte row kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c15
echo "== kernel + model tests"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_seq_attention_gpu.py -q -m gpu -x 2>&1 | tail -3 | cut -c1-300
for rep in 1 2; do
echo "== resident bench, overlaps on (rep $rep)"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
echo "== resident bench, overlaps off (rep $rep)"; BUGLAB_B200_OVERLAP=0 timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
done
echo "== bench_seq"; timeout 600 python scripts/bench_seq.py --steps 8 > ${O}_bench_seq.json 2> ${O}_bench_seq.err; echo "rc=$?"; tail -1 ${O}_bench_seq.json | cut -c1-700
echo "== smoke"; timeout 900 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-300
