#!/bin/bash
# round 2, fourteenth GPU pass: the producers' wave barrier -- correctness, DRAM traffic and sustained time per period
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2n_$name.log 2> $O/r2n_$name.err; echo "$name exit $?" >> $O/r2n_summary.txt; }
: > $O/r2n_summary.txt
ZETT_GEMM_WAVE_SYNC=1 run gemm5_sync1 600 python tests/gpu_selftest.py gemm --impl 5
ZETT_GEMM_WAVE_SYNC=1 run gemm2_sync1 600 python tests/gpu_selftest.py gemm --impl 2
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
for sy in 0 1 2 4 8 16; do
  ZETT_GEMM_WAVE_SYNC=$sy run ncu_sync$sy 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --mnk 53248,12288,4096
  ZETT_GEMM_WAVE_SYNC=$sy run raster_sync$sy 300 python tests/gpu_selftest.py raster --mnk 53248,12288,4096 --combos "48,4,2,2;48,4,2,3"
done
run ncu_halfm 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --mnk 26624,12288,4096
for sy in 0 1 4; do
  ZETT_GEMM_WAVE_SYNC=$sy run ncu256_sync$sy 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --impl 2 --mnk 53248,12288,4096
  ZETT_GEMM_WAVE_SYNC=$sy run bench_sync$sy 600 python bench.py --no-cpu-baseline --no-extra
  ZETT_GEMM_WAVE_SYNC=$sy run bench_xlmr_sync$sy 600 python bench.py --config xlmr --no-cpu-baseline --no-extra
done
