#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt gpurun_out/raster_*.csv
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,lts__t_sector_op_read_hit_rate.pct"
for cfg in "12 2" "12 4" "24 2" "24 4" "32 4" "48 4" "48 8" "100000 16"; do
  set -- $cfg
  ZETT_RASTER_CHUNK_MB=$1 ZETT_RASTER_GROUP_M=$2 timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv \
    --log-file gpurun_out/raster_$1_$2.csv python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms 2 > /dev/null 2>&1
  echo "raster $cfg exit $?" >> gpurun_out/summary.txt
done
for cfg in "24 4" "100000 16"; do
  set -- $cfg
  ZETT_RASTER_CHUNK_MB=$1 ZETT_RASTER_GROUP_M=$2 timeout 600 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -s 1 -c 1 --csv \
    --log-file gpurun_out/raster_small_$1_$2.csv python tests/gpu_selftest.py one --mnk 16384,4096,4096 --impl 2 --terms 2 > /dev/null 2>&1
done
cat gpurun_out/summary.txt
