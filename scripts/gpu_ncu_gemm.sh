#!/bin/bash
# ncu --set full of the dominant kernel in its default configuration (fp16 + 2 x e5m2 operands, 256 x 512 pair tiles):
# one isolated launch of the QKV GEMM of a 53 248-position pass.  Read back with scripts/ncu_summary.py.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
timeout 280 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/gemm_${R}_f16f8_wide -f \
  python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 5 --terms 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
tail -n 3 gpurun_out/ncu_full.log
