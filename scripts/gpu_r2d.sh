#!/bin/bash
# round 2, fourth GPU pass: relaxed accumulator-release arrive + predicated-store quads; launch lists and ncu --set full
# captures of every hot kernel (profiles/*_r2_*)
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2d_$name.log 2> $O/r2d_$name.err; echo "$name exit $?" >> $O/r2d_summary.txt; }
: > $O/r2d_summary.txt
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run gemm2 600 python tests/gpu_selftest.py gemm --impl 2
if grep -q '"ok": false\|error' $O/r2d_gemm2.log $O/r2d_gemm5.log; then echo "GEMM FAILED" >> $O/r2d_summary.txt; exit 0; fi
run fwd_big 1200 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,xlmr,tinyllama,mistral
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2 --mnk "53248,12288,4096;54000,2304,768;54000,768,1536;53248,6144,2048"
ZETT_SUSTAINED_ONLY="f16+2xe5m2 256x256,f16+2xe5m2 256x512" run sustained 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096"
run bench_full 1500 python bench.py
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
# launch lists of one 16384-row pass (cold-cache, serialised: compare SHARES)
run ncu_launches_mistral 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 140 --csv --log-file $O/launches_r2_mistral.csv python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_launches_xlmr 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 140 --csv --log-file $O/launches_r2_xlmr.csv python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
# --set full of the dominant kernel in both tile shapes, and of the HBM-bound kernels of a Mistral pass
run ncu_gemm_wide 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2_f16f8_wide -f python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 5 --terms 2
run ncu_gemm_256 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2_f16f8 -f python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms 2
run ncu_gemm_xlmr 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o $O/gemm_r2_f16f8_xlmr -f python tests/gpu_selftest.py one --mnk 54000,2304,768 --impl 5 --terms 2
run ncu_ln 900 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 20 -c 3 -o $O/layernorm_r2 -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_attn 900 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 6 -c 2 -o $O/attention_r2 -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
run ncu_gather 900 ncu --set full --clock-control none --import-source on -k regex:gather_rescale -s 2 -c 1 -o $O/gather_r2 -f python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline --no-extra --parity-rows 8
