#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/bk_*.csv
for impl in 1 2; do
  ZETT_BLOCK_K=32 timeout 600 python tests/gpu_selftest.py gemm --impl $impl > gpurun_out/bk32_gemm_impl$impl.log 2>&1
  echo "bk32 gemm impl $impl exit $?" >> gpurun_out/summary.txt
done
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum"
for bk in 64 32; do
 for terms in 2 3 1; do
  ZETT_BLOCK_K=$bk timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv \
    --log-file gpurun_out/bk_${bk}_t${terms}.csv python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms $terms > gpurun_out/bk_run.log 2>&1
  echo "bk $bk terms $terms exit $?" >> gpurun_out/summary.txt
 done
done
ZETT_BLOCK_K=32 timeout 900 python bench.py --steps 5 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/f8_bench_mistral_bk32.log 2>&1
echo "bench mistral terms2 bk32 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/f8_bench_mistral.log 2>&1
echo "bench mistral terms2 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mistral_terms3.log 2>&1
echo "bench mistral terms3 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
