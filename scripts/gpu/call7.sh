#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c7
echo "== smoke gradient diagnosis (400-node graphs)"; timeout 900 python scripts/diag_smoke_grads.py 256 400 > ${O}_diag_400.txt 2>&1; echo "rc=$?"; cut -c1-1400 ${O}_diag_400.txt
echo "== same at 2000-node graphs"; timeout 900 python scripts/diag_smoke_grads.py 256 2000 > ${O}_diag_2000.txt 2>&1; echo "rc=$?"; cut -c1-900 ${O}_diag_2000.txt
