#!/bin/bash
# round 2, sixteenth GPU pass: with the wave barrier on, does any rasterisation / hint combination keep the W chunk in L2?
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2p_$name.log 2> $O/r2p_$name.err; echo "$name exit $?" >> $O/r2p_summary.txt; }
: > $O/r2p_summary.txt
COMBOS="48,4,2,2;48,4,2,3;48,4,1,3;32,4,2,3;32,9,2,3;24,4,2,3;64,4,2,3;48,12,2,3;48,2,2,3;96,4,2,3"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
IFS=';' read -ra CS <<< "$COMBOS"
for c in "${CS[@]}"; do
  IFS=',' read -r C G A W <<< "$c"
  ZETT_RASTER_CHUNK_MB=$C ZETT_RASTER_GROUP_M=$G ZETT_L2_HINT_A=$A ZETT_L2_HINT_W=$W run ncu_${C}_${G}_${A}_${W} 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --mnk 53248,12288,4096
done
run raster 600 python tests/gpu_selftest.py raster --mnk 53248,12288,4096 --combos "$COMBOS"
