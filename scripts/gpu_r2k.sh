#!/bin/bash
# round 2, eleventh GPU pass: (1) rasterisation / L2-hint probes of the QKV GEMM -- sustained time and, per combination,
# DRAM / L2-fabric traffic under ncu; (2) attention with K/V prefetch and four resident blocks
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2k_$name.log 2> $O/r2k_$name.err; echo "$name exit $?" >> $O/r2k_summary.txt; }
: > $O/r2k_summary.txt
COMBOS="48,4,2,2;48,4,1,3;48,4,2,3;24,4,1,3;24,8,1,3;16,12,1,3;96,4,2,2;24,4,2,2;48,12,2,2;48,12,1,3;32,9,1,3;200,3,2,2;48,6,1,3;48,2,2,2"
run raster 600 python tests/gpu_selftest.py raster --mnk 53248,12288,4096 --combos "$COMBOS"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
IFS=';' read -ra CS <<< "$COMBOS"
for c in "${CS[@]}"; do
  IFS=',' read -r C G A W <<< "$c"
  ZETT_RASTER_CHUNK_MB=$C ZETT_RASTER_GROUP_M=$G ZETT_L2_HINT_A=$A ZETT_L2_HINT_W=$W run ncu_${C}_${G}_${A}_${W} 300 ncu --metrics $M --clock-control none -k regex:gemm_tcgen05 -c 1 --csv python tests/gpu_selftest.py one --mnk 53248,12288,4096
done
for pf in 0 1; do
  ZETT_ATTN_PREFETCH=$pf run bench_xlmr_pf$pf 600 python bench.py --config xlmr --no-cpu-baseline --no-extra
  ZETT_ATTN_PREFETCH=$pf run bench_mistral_pf$pf 600 python bench.py --no-cpu-baseline --no-extra
  ZETT_ATTN_PREFETCH=$pf run bench_tinyllama_pf$pf 600 python bench.py --config tinyllama --no-cpu-baseline --no-extra
done
ZETT_ATTN_PREFETCH=1 run fwd 900 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,xlmr,mistral
