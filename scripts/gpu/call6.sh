#!/usr/bin/env bash
# Round-2 GPU call 6 (2 GPUs): data-parallel bench with the overlapped bucketed all-reduce, reference arm with a GPU visible,
# smoke(), kernel + model tests of the two-rows-per-warp by-source edge backward.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c6
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; echo "rc=$?"; tail -3 ${O}_smoke.txt | cut -c1-400
echo "== kernel + model tests"; timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x > ${O}_tests.txt 2>&1; echo "rc=$?"; tail -4 ${O}_tests.txt
echo "== 2-GPU bench"; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err; echo "rc=$?"; cut -c1-1500 ${O}_bench_2gpu.json; tail -3 ${O}_bench_2gpu.err | cut -c1-300
echo "== reference arm, GPUs visible, under torchrun"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "rc=$?"; cut -c1-500 ${O}_bench_reference.json; tail -2 ${O}_bench_reference.err | cut -c1-300
echo "== 1-GPU bench, resident only"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline > ${O}_bench_1gpu.json 2>/dev/null; echo "rc=$?"; cut -c1-600 ${O}_bench_1gpu.json
