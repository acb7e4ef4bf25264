#!/bin/bash
# round 2, first GPU pass over the rewritten GEMM engine: correctness first, then stall pictures and sustained probes
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2a_gpu.txt 2>&1
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  timeout $t "$@" > $O/r2a_$name.log 2> $O/r2a_$name.err
  echo "$name exit $?" >> $O/r2a_summary.txt
}
: > $O/r2a_summary.txt
run gemm3 600 python tests/gpu_selftest.py gemm --impl 3
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
if grep -q '"ok": false\|error' $O/r2a_gemm2.log $O/r2a_gemm5.log; then
  echo "GEMM FAILED - diagnostics only" >> $O/r2a_summary.txt
  ZETT_GEMM_PROF=1 run one_small 120 python tests/gpu_selftest.py one --mnk 512,256,128 --impl 2 --terms 2
  exit 0
fi
run fwd_tiny5 600 python tests/gpu_selftest.py forward --impl 5
run fwd_tiny3 600 python tests/gpu_selftest.py forward --impl 3
run fwd_tiny5_t3 600 python tests/gpu_selftest.py forward --impl 5 --terms 3
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 900 python tests/gpu_selftest.py sweep
run sustained 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096;16384,4096,8192"
run bench_mistral 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
run bench_xlmr 600 python bench.py --config xlmr --steps 5 --warmup 3 --no-cpu-baseline
run bench_tinyllama 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline
