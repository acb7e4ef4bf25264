#!/bin/bash
# iteration run: sweep + f8 parity + bench in both precisions
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py sweep > gpurun_out/sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/summary.txt
for impl in 3 2; do
  timeout 600 python tests/gpu_selftest.py gemm --impl $impl > gpurun_out/gemm_impl$impl.log 2>&1
  echo "gemm impl $impl exit $?" >> gpurun_out/summary.txt
done
for impl in 3 1 2; do
  timeout 900 python tests/gpu_selftest.py forward --impl $impl --terms 2 > gpurun_out/f8_fwd_impl$impl.log 2>&1
  echo "forward terms2 impl $impl exit $?" >> gpurun_out/summary.txt
done
timeout 900 python tests/gpu_selftest.py forward --impl 2 --terms 3 > gpurun_out/fwd_impl2.log 2>&1
echo "forward terms3 impl 2 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 2 --terms 2 --configs xlmr,tinyllama,mistral > gpurun_out/f8_fwd_big.log 2>&1
echo "forward big terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 2 --terms 3 --configs xlmr,tinyllama,mistral > gpurun_out/fwd_big.log 2>&1
echo "forward big terms3 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/f8_bench_mistral.log 2>&1
echo "bench mistral terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mistral_terms3.log 2>&1
echo "bench mistral terms3 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
