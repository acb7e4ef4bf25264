#!/bin/bash
# (1) interleaved fp8 planes: GEMM unit tests (tcgen05 pairs, pairs of pairs, SIMT) + forward parity in that format;
# (2) sustained energy decomposition (loads / MMA terms / stores, L2 eviction hints, streaming stores);
# (3) ncu DRAM / L2 bytes with and without the hints; (4) bench in the candidate configurations.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for impl in 2 3 4; do
  timeout 600 python tests/gpu_selftest.py gemm --impl $impl > gpurun_out/p3_gemm_impl$impl.log 2>&1
  echo "gemm impl $impl exit $?" >> gpurun_out/summary.txt
done
for impl in 2 4; do
  timeout 600 python tests/gpu_selftest.py forward --impl $impl --terms 2 > gpurun_out/p3_fwd_impl${impl}_t2.log 2>&1
  echo "forward impl $impl terms 2 exit $?" >> gpurun_out/summary.txt
done
timeout 900 python tests/gpu_selftest.py forward --impl 2 --terms 2 --configs xlmr,tinyllama,mistral > gpurun_out/p3_fwd_big_t2.log 2>&1
echo "forward big terms 2 exit $?" >> gpurun_out/summary.txt
timeout 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096" > gpurun_out/p3_sustained.log 2>&1
echo "sustained exit $?" >> gpurun_out/summary.txt
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second"
i=0
for v in "3 2 2 0" "3 3 2 1" "3 3 1 1" "2 2 2 0" "2 3 2 1"; do
  set -- $v
  ZETT_L2_HINT_W=$2 ZETT_L2_HINT_A=$3 ZETT_STREAM_OUT=$4 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv \
    --log-file gpurun_out/p3_ncu_t$1_w$2_a$3_s$4.csv python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms $1 > /dev/null 2>&1
  echo "ncu terms $1 hintW $2 hintA $3 stream $4 exit $?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 3 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/p3_bench_t2.log 2>&1
echo "bench terms 2 exit $?" >> gpurun_out/summary.txt
ZETT_L2_HINT_W=3 ZETT_STREAM_OUT=1 timeout 600 python bench.py --steps 3 --warmup 3 --split-terms 2 --no-cpu-baseline > gpurun_out/p3_bench_t2_hints.log 2>&1
echo "bench terms 2 hints exit $?" >> gpurun_out/summary.txt
ZETT_L2_HINT_W=3 ZETT_STREAM_OUT=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p3_bench_t3_hints.log 2>&1
echo "bench terms 3 hints exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
