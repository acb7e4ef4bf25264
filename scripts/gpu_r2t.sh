#!/bin/bash
# round 2, last GPU pass: the wave barrier's give-up flag -- GEMM self-tests, the driver's test command, the default bench
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2t_$name.log 2> $O/r2t_$name.err; echo "$name exit $?" >> $O/r2t_summary.txt; }
: > $O/r2t_summary.txt
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
run bench_full 1500 python bench.py
ZETT_SUSTAINED_ONLY="f16+2xe5m2 256x512" run sustained 300 python tests/gpu_selftest.py sustained --mnk 53248,12288,4096
