#!/bin/bash
# round 2: the default bench line of the final tree (first-pass split only when the retokenizer is short of threads)
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/r2u_bench_full.log 2> $O/r2u_bench_full.err; echo "bench exit $?" > $O/r2u_summary.txt
timeout 300 python -m pytest tests/test_gpu_native.py -x -q -m gpu -k "pipelined or module_matches" > $O/r2u_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2u_summary.txt
