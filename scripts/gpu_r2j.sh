#!/bin/bash
# round 2, tenth GPU pass: LayerNorm / attention instantiated per operand format, attention with eight lanes per head;
# launch lists and ncu captures of both on the new build; the driver's test command
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2j_$name.log 2> $O/r2j_$name.err; echo "$name exit $?" >> $O/r2j_summary.txt; }
: > $O/r2j_summary.txt
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,tiny_single_head,tiny_plain,tiny_one_layer,xlmr,tinyllama,mistral
run fwd_t3 900 python tests/gpu_selftest.py forward --impl 0 --terms 3 --configs tiny,xlmr
run bench_full 1500 python bench.py
run ncu_launches_mistral 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 90 --csv --log-file $O/launches_r2j_mistral.csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_xlmr 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 100 --csv --log-file $O/launches_r2j_xlmr.csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_ln 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 20 -c 3 -o $O/layernorm_r2j -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_ln_xlmr 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 24 -c 3 -o $O/layernorm_r2j_xlmr -f python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_attn 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 6 -c 2 -o $O/attention_r2j -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_attn_xlmr 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 8 -c 2 -o $O/attention_r2j_xlmr -f python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
