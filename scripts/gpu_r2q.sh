#!/bin/bash
# round 2, seventeenth GPU pass: LayerNorm with per-row cursors (64 registers / 80 registers), parity + per-launch times
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2q_$name.log 2> $O/r2q_$name.err; echo "$name exit $?" >> $O/r2q_summary.txt; }
: > $O/r2q_summary.txt
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,tiny_single_head,tiny_plain,tiny_one_layer,tiny_multi_pass,xlmr,tinyllama,mistral
run fwd_t3 600 python tests/gpu_selftest.py forward --impl 0 --terms 3 --configs tiny,xlmr
ZETT_LN_MINB3=1 run fwd_minb3 900 python tests/gpu_selftest.py forward --impl 0 --configs tiny,xlmr,mistral
M=gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
for v in 4 3; do
  if [ $v = 3 ]; then export ZETT_LN_MINB3=1; else unset ZETT_LN_MINB3; fi
  run ncu_ln_mistral_b$v 600 ncu --metrics $M --clock-control none -k regex:layernorm_kernel -s 20 -c 8 --csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
  run ncu_ln_xlmr_b$v 600 ncu --metrics $M --clock-control none -k regex:layernorm_kernel -s 24 -c 10 --csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
  run bench_mistral_b$v 600 python bench.py --no-cpu-baseline --no-extra
  run bench_xlmr_b$v 600 python bench.py --config xlmr --no-cpu-baseline --no-extra
  run bench_tinyllama_b$v 600 python bench.py --config tinyllama --no-cpu-baseline --no-extra
done
