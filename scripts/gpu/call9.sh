#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c9
S=scripts/diag_smoke_step.py
BUGLAB_B200_TMA=0 timeout 300 python $S /tmp/g_notma.pt 2>&1 | grep -v Warn | tail -3
timeout 300 python $S /tmp/g_tma.pt /tmp/g_notma.pt 2>&1 | grep -v Warn | tail -4
timeout 300 python $S /tmp/g_tma2.pt /tmp/g_notma.pt /tmp/g_tma.pt 2>&1 | grep -v Warn | tail -4
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 python $S /tmp/g_tma_nocache.pt /tmp/g_notma.pt /tmp/g_tma.pt 2>&1 | grep -v Warn | tail -4
CUDA_LAUNCH_BLOCKING=1 timeout 600 python $S /tmp/g_tma_blocking.pt /tmp/g_notma.pt /tmp/g_tma.pt 2>&1 | grep -v Warn | tail -4
BUGLAB_B200_OVERLAP=0 timeout 600 python $S /tmp/g_tma_noov.pt /tmp/g_notma.pt 2>&1 | grep -v Warn | tail -4
echo "== initcheck"
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 1500 compute-sanitizer --tool initcheck --print-limit 30 python $S /tmp/g_init.pt /tmp/g_notma.pt > ${O}_initcheck.txt 2>&1; echo "rc=$?"
grep -c "Uninitialized" ${O}_initcheck.txt; grep -A12 "Uninitialized" ${O}_initcheck.txt | grep -v "^=========     Host Frame.*\(libtorch\|libc10\|python\|libcuda\)" | head -80
echo "== memcheck"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 30 python $S /tmp/g_mem.pt /tmp/g_notma.pt > ${O}_memcheck.txt 2>&1; echo "rc=$?"
grep -B2 -A14 "Invalid\|out of bounds" ${O}_memcheck.txt | head -60; tail -5 ${O}_memcheck.txt
