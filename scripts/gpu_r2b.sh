#!/bin/bash
# round 2, second GPU pass: lean epilogue (rolled quads, single register chunk), comm exports
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2b_$name.log 2> $O/r2b_$name.err; echo "$name exit $?" >> $O/r2b_summary.txt; }
: > $O/r2b_summary.txt
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
if grep -q '"ok": false\|error' $O/r2b_gemm2.log $O/r2b_gemm5.log; then echo "GEMM FAILED" >> $O/r2b_summary.txt; exit 0; fi
run fwd_tiny5 600 python tests/gpu_selftest.py forward --impl 5
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2,1 --mnk "53248,12288,4096;53248,8192,4096;54000,2304,768;54000,1536,768;54000,768,1536;53248,6144,2048;53248,4096,2048;16384,4096,4096"
ZETT_SUSTAINED_ONLY="f16+2xe5m2 256x256,f16+2xe5m2 256x512,bf16 single pass 256x256" run sustained 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096"
run bench_mistral 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
run bench_xlmr 600 python bench.py --config xlmr --steps 5 --warmup 3 --no-cpu-baseline
run bench_tinyllama 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline
ZETT_WIDE_MIN_K=2048 run bench_tinyllama_wide2048 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline
run pytest_comm 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu
