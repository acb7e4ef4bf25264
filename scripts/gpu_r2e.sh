#!/bin/bash
# round 2, fifth GPU pass: LayerNorm occupancy fix; launch lists of one pass; the driver's test command
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2e_$name.log 2> $O/r2e_$name.err; echo "$name exit $?" >> $O/r2e_summary.txt; }
: > $O/r2e_summary.txt
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,xlmr,tinyllama,mistral
run bench_full 1500 python bench.py
# launch lists of one 16384-row pass (cold-cache, serialised: compare SHARES): 27 weight-split launches, 42 per pass
run ncu_launches_mistral 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 90 --csv --log-file $O/launches_r2_mistral.csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_xlmr 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 100 --csv --log-file $O/launches_r2_xlmr.csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_ln 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 20 -c 3 -o $O/layernorm_r2 -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
