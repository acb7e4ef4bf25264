#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
i=0
for cfg in "53248,12288,4096 2" "53248,12288,4096 3" "16384,4096,4096 2" "16384,4096,4096 3"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -f -o gpurun_out/one_$i \
    python tests/gpu_selftest.py one --mnk $1 --impl 2 --terms $2 > gpurun_out/one_$i.log 2>&1
  echo "ncu one $cfg exit $?" >> gpurun_out/summary.txt
  i=$((i+1))
done
cat gpurun_out/summary.txt
