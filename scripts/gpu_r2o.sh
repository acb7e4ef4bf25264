#!/bin/bash
# round 2, fifteenth GPU pass: wave barrier on by default -- GEMM self-tests, sanitizers, W-hint A/B, the driver's test
# command, the default bench, sustained probes and fresh ncu --set full captures of the GEMM shapes
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2o_$name.log 2> $O/r2o_$name.err; echo "$name exit $?" >> $O/r2o_summary.txt; }
: > $O/r2o_summary.txt
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
if grep -q '"ok": false' $O/r2o_gemm5.log $O/r2o_gemm2.log; then echo "GEMM FAILED" >> $O/r2o_summary.txt; exit 0; fi
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,tiny_multi_pass,xlmr,tinyllama,mistral
run memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang
run memcheck_gemm 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_selftest.py one --mnk 8192,2048,512
run racecheck 900 compute-sanitizer --tool racecheck --racecheck-report all python tests/gpu_selftest.py forward --impl 0 --configs tiny
run bench_whint_normal 600 python bench.py --no-cpu-baseline --no-extra
ZETT_L2_HINT_W=3 run bench_whint_last 600 python bench.py --no-cpu-baseline --no-extra
ZETT_GEMM_WAVE_SYNC=0 run bench_nosync 600 python bench.py --no-cpu-baseline --no-extra
run pytest_gpu 1800 python -m pytest tests -x -q -m gpu
run bench_full 1500 python bench.py
ZETT_SUSTAINED_ONLY="256x512,single pass 256x256" run sustained 600 python tests/gpu_selftest.py sustained --mnk 53248,12288,4096
run ncu_gemm_wide 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2o_f16f8_wide -f python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 5 --terms 2
run ncu_gemm_256 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2o_f16f8 -f python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms 2
run ncu_gemm_xlmr 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2o_f16f8_xlmr -f python tests/gpu_selftest.py one --mnk 54000,2304,768 --impl 5 --terms 2
run ncu_launches_mistral 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 90 --csv --log-file $O/launches_r2o_mistral.csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_xlmr 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 100 --csv --log-file $O/launches_r2o_xlmr.csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
