#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c8
timeout 900 python scripts/diag_smoke_layers.py 256 400 > ${O}_layers_400.txt 2>&1; echo "rc=$?"; cut -c1-260 ${O}_layers_400.txt | grep -v Warning | head -80
