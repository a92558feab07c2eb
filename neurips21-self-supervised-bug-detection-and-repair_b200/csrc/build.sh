#!/usr/bin/env bash
# Builds libbuglab_b200.so in-tree for sm_100a (the only target).  Usage: csrc/build.sh [-v]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../buglab_b200/libbuglab_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr)
if [[ "${1:-}" == "-v" ]]; then FLAGS+=(-Xptxas -v); fi
SRCS=(api.cu plan.cu rows.cu edge_segmax.cu gemm.cu pair_project_tc.cu gemm_tma.cu node_update.cu segment_ops.cu embed.cu optim.cu seq_attention.cu seq_attention_tc.cu)
OBJS=()
pids=()
mkdir -p build
for s in "${SRCS[@]}"; do
  o="build/${s%.cu}.o"
  OBJS+=("$o")
  if [[ ! -f "$o" || "$s" -nt "$o" || common.cuh -nt "$o" || seq_attention_core.h -nt "$o" || ../../include/buglab_b200.h -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT" "${OBJS[@]}" -lcudart -lcublas
echo "built $OUT"
# host-only shard decoder (no CUDA): include/buglab_shards.h
SHARDS_OUT=../buglab_b200/libbuglab_shards.so
if [[ ! -f "$SHARDS_OUT" || shards/shards.cpp -nt "$SHARDS_OUT" || ../../include/buglab_shards.h -nt "$SHARDS_OUT" ]]; then
  ${CXX:-g++} -O3 -std=c++17 -fPIC -shared -Wall -Wextra -o "$SHARDS_OUT" shards/shards.cpp -lz
fi
echo "built $SHARDS_OUT"
