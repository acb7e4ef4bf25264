#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1200 python tests/gpu_selftest.py sweep > gpurun_out/sweep2.log 2>&1
echo "sweep exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py gemm --impl 2 > gpurun_out/gemm_impl2.log 2>&1
echo "gemm impl 2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/f8_bench_mistral.log 2>&1
echo "bench mistral terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mistral_terms3.log 2>&1
echo "bench mistral terms3 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
