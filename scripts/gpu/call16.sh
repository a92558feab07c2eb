#!/usr/bin/env bash
# Round-2 GPU call 16 (2 GPUs): data-parallel step with 32 MB overlapped buckets vs one bucket after backward (round 1's scheme).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c16
run() { # label, extra env
  echo "== $1"; env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>${O}_$3.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['allreduce']['buckets'], d['allreduce']['exposed_ms_last_step'], d['clocks']['sm_mhz'])"
}
echo "== 1 GPU"; timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks']['sm_mhz'])"
run "2 GPUs, 32 MB buckets (default)" "X=1" 29521
run "2 GPUs, one bucket after backward" "BUGLAB_B200_ALLREDUCE_BUCKET_MB=4096" 29522
run "2 GPUs, 8 MB buckets" "BUGLAB_B200_ALLREDUCE_BUCKET_MB=8" 29523
run "2 GPUs, 32 MB buckets again" "X=1" 29524
