#!/bin/bash
# round 2, twelfth GPU pass: which die each SM is on, and whether data read from both dies is fetched from DRAM twice
set -u
O=gpurun_out
mkdir -p $O
nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_die_probe scripts/l2_die_probe.cu > $O/r2l_build.log 2>&1 || exit 1
timeout 120 /tmp/l2_die_probe > $O/r2l_probe.log 2>&1
timeout 120 /tmp/l2_die_probe > $O/r2l_probe_again.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum --clock-control none -k regex:shared_read --csv /tmp/l2_die_probe > $O/r2l_ncu.log 2>&1
