#!/bin/bash
# fp16 + e5m2-correction mode: GEMM unit tests per implementation, forward parity, bench; plus the ncu launch list.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
rm -f gpurun_out/summary.txt
for impl in 3 1 2; do
  timeout 600 python tests/gpu_selftest.py gemm --impl $impl > gpurun_out/f8_gemm_impl$impl.log 2>&1
  echo "gemm impl $impl exit $?" >> gpurun_out/summary.txt
done
for impl in 3 1 2; do
  timeout 900 python tests/gpu_selftest.py forward --impl $impl --terms 2 > gpurun_out/f8_fwd_impl$impl.log 2>&1
  echo "forward terms2 impl $impl exit $?" >> gpurun_out/summary.txt
done
timeout 900 python tests/gpu_selftest.py forward --impl 2 --terms 2 --configs xlmr,tinyllama,mistral > gpurun_out/f8_fwd_big.log 2>&1
echo "forward big terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/f8_bench_mistral.log 2>&1
echo "bench mistral terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mistral_terms3.log 2>&1
echo "bench mistral terms3 exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
