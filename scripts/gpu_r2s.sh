#!/bin/bash
# round 2, final single-GPU pass: first-pass split A/B under two retokenizer threads, the driver's commands (pytest -m gpu,
# smoke, bench, reference arm), LayerNorm captures and launch lists of the final build
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2s_$name.log 2> $O/r2s_$name.err; echo "$name exit $?" >> $O/r2s_summary.txt; }
: > $O/r2s_summary.txt
LOCAL_WORLD_SIZE=8 ZETT_BENCH_FIRST_CHUNK=0 run bench_2threads_nosplit 600 python bench.py --no-cpu-baseline --no-extra
LOCAL_WORLD_SIZE=8 ZETT_BENCH_FIRST_CHUNK=4096 run bench_2threads_split 600 python bench.py --no-cpu-baseline --no-extra
ZETT_BENCH_FIRST_CHUNK=0 run bench_16threads_nosplit 600 python bench.py --no-cpu-baseline --no-extra
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
run smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
run bench_full 1500 python bench.py
run bench_reference 900 python bench.py --impl reference --steps 3 --warmup 1
run ncu_ln 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 20 -c 3 -o $O/layernorm_r2s -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_ln_xlmr 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 24 -c 3 -o $O/layernorm_r2s_xlmr -f python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_mistral 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 90 --csv --log-file $O/launches_r2s_mistral.csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_xlmr 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 100 --csv --log-file $O/launches_r2s_xlmr.csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
