#!/bin/bash
# round 2, seventh GPU pass: branch-free GELUs, register-resident attention; full GPU test suite
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2g_$name.log 2> $O/r2g_$name.err; echo "$name exit $?" >> $O/r2g_summary.txt; }
: > $O/r2g_summary.txt
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run gemm3 600 python tests/gpu_selftest.py gemm --impl 3
if grep -q '"ok": false\|error' $O/r2g_gemm5.log; then echo "GEMM FAILED" >> $O/r2g_summary.txt; exit 0; fi
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,tiny_single_head,tiny_plain,tiny_one_layer,xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2 --mnk "53248,8192,4096;54000,1536,768"
run bench_full 1500 python bench.py
ZETT_ATTN_STREAMING=1 run bench_xlmr_attn_streaming 600 python bench.py --config xlmr --no-cpu-baseline --no-extra
run bench_xlmr 600 python bench.py --config xlmr --no-cpu-baseline --no-extra
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
