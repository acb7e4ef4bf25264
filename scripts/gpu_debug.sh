#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
ZETT_DEBUG=1 timeout 600 python tests/gpu_selftest.py forward --impl 3 --terms 2 --configs tiny > gpurun_out/dbg_fwd_tiny_f8_simt.log 2>&1
echo "dbg tiny f8 simt exit $?" >> gpurun_out/summary.txt
ZETT_DEBUG=1 timeout 600 python tests/gpu_selftest.py forward --impl 3 --terms 3 --configs tiny > gpurun_out/dbg_fwd_tiny_bf16_simt.log 2>&1
echo "dbg tiny bf16 simt exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py sweep > gpurun_out/sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
