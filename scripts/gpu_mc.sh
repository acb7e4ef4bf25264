#!/bin/bash
# Cluster-of-four W-multicast GEMM (gemm_impl 4): unit tests, forward parity, timing sweep against CTA pairs (impl 2) in
# both operand formats, full bench in the four combinations.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
export ZETT_VERBOSE=1
timeout 600 python tests/gpu_selftest.py gemm --impl 4 > gpurun_out/mc_gemm_impl4.log 2>&1
echo "gemm impl 4 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py forward --impl 4 > gpurun_out/mc_fwd_impl4.log 2>&1
echo "forward impl 4 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py forward --impl 4 --terms 2 > gpurun_out/mc_fwd_impl4_t2.log 2>&1
echo "forward impl 4 terms 2 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 4 --terms 2 --configs xlmr,tinyllama,mistral > gpurun_out/mc_fwd_big_t2.log 2>&1
echo "forward big impl 4 terms 2 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py sweep > gpurun_out/mc_sweep.log 2>&1
echo "sweep exit $?" >> gpurun_out/summary.txt
for combo in "4 2" "4 3" "2 2" "2 3"; do
  set -- $combo
  timeout 600 python bench.py --steps 3 --warmup 3 --gemm-impl $1 --split-terms $2 --no-cpu-baseline > gpurun_out/mc_bench_i$1_t$2.log 2>&1
  echo "bench impl $1 terms $2 exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
grep -h '"kind": "sweep"' gpurun_out/mc_sweep.log | cut -c1-200
grep -h "co-resident" gpurun_out/*.log | sort | uniq -c
for f in gpurun_out/mc_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-330; done
