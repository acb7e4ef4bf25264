#!/bin/bash
# round 2, third GPU pass: instruction-lean epilogue with prefetched column constants, rewritten bench (parity, extras)
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2c_$name.log 2> $O/r2c_$name.err; echo "$name exit $?" >> $O/r2c_summary.txt; }
: > $O/r2c_summary.txt
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run gemm3 600 python tests/gpu_selftest.py gemm --impl 3
if grep -q '"ok": false\|error' $O/r2c_gemm2.log $O/r2c_gemm5.log; then echo "GEMM FAILED" >> $O/r2c_summary.txt; exit 0; fi
run fwd_tiny5 600 python tests/gpu_selftest.py forward --impl 5
run fwd_tiny5_t3 600 python tests/gpu_selftest.py forward --impl 5 --terms 3
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2 --mnk "53248,12288,4096;53248,8192,4096;54000,2304,768;54000,1536,768;54000,768,1536;53248,6144,2048;53248,4096,2048"
ZETT_SUSTAINED_ONLY="f16+2xe5m2 256x256,f16+2xe5m2 256x512,bf16 single pass 256x256" run sustained 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096"
run bench_full 1500 python bench.py
ZETT_WIDE_MIN_K=2048 run bench_tinyllama_wide2048 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline --no-extra
run bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
