#!/bin/bash
# round 2, sixth GPU pass: straight-line epilogue instantiations + residual prefetch; 2-GPU sharded path is a separate call
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2f_$name.log 2> $O/r2f_$name.err; echo "$name exit $?" >> $O/r2f_summary.txt; }
: > $O/r2f_summary.txt
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
if grep -q '"ok": false\|error' $O/r2f_gemm2.log $O/r2f_gemm5.log; then echo "GEMM FAILED" >> $O/r2f_summary.txt; exit 0; fi
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,tiny_single_head,tiny_plain,xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2 --mnk "53248,4096,4096;54000,768,768;54000,2304,768;54000,768,1536;53248,2048,2048"
run bench_full 1500 python bench.py
run pytest_guard 900 python -m pytest tests/test_gpu_native.py -x -q -m gpu -k "guard or masked"
