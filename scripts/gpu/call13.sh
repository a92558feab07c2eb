#!/usr/bin/env bash
# Round-2 GPU call 13: ncu --set full of the tensor-core attention's kernels at the config-4 shape; seq tests incl. the
# full-size fp64-oracle cases; bench_seq after the LayerNorm-reduction fix.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c13
echo "== seq tests"; timeout 1500 python -m pytest tests/test_seq_attention_gpu.py tests/test_kernels_gpu.py -q -m gpu 2>&1 | tail -6 | cut -c1-400
echo "== bench_seq"; timeout 600 python scripts/bench_seq.py --steps 8 > ${O}_bench_seq.json 2> ${O}_bench_seq.err; echo "rc=$?"; tail -1 ${O}_bench_seq.json | cut -c1-700
echo "== ncu attention"; timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"softmax|proj_kernel|wgrad_kernel" --launch-skip 8 -c 8 -f -o ${O}_seq_attention python scripts/profile_seq_attention.py > ${O}_ncu.log 2>&1; echo "rc=$?"; tail -3 ${O}_ncu.log | cut -c1-200
python scripts/ncu_summary.py ${O}_seq_attention.ncu-rep "tensor-core attention at the config-4 shape (64 x 8 heads x 512 tokens)" > ${O}_seq_attention_ncu.md 2>/dev/null; grep -c "^## " ${O}_seq_attention_ncu.md
