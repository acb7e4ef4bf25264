#!/bin/bash
# round 2, thirteenth GPU pass: how the GEMM's DRAM traffic grows with the number of CTA pairs that share an A tile
set -u
O=gpurun_out
mkdir -p $O
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
for spec in "512 4" "1024 4" "1024 1" "1536 4" "3072 4" "3072 1" "3072 12" "6144 4"; do
  set -- $spec
  ZETT_RASTER_GROUP_M=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --mnk 53248,$1,4096 > $O/r2m_ncu_n$1_g$2.log 2> $O/r2m_ncu_n$1_g$2.err
done
