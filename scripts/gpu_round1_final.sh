#!/bin/bash
# Round gate: the driver's GPU test command, smoke(), bench at several pass sizes, the ncu launch list of one pass and
# a --set full capture of the dominant kernel (largest launch) in the default operand format.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/f_pytest_gpu.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
for rpp in 0 25152 50304; do
  timeout 600 python bench.py --steps 3 --warmup 3 --rows-per-pass $rpp --no-cpu-baseline > gpurun_out/f_bench_rpp$rpp.log 2>&1
  echo "bench rows-per-pass $rpp exit $?" >> gpurun_out/summary.txt
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 1 -c 1 -o gpurun_out/gemm_${R}_f16f8 -f \
  python tests/gpu_selftest.py one --mnk 53248,12288,4096 --impl 2 --terms 2 > gpurun_out/f_ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/f_pytest_gpu.log
tail -n 2 gpurun_out/f_smoke.log
for f in gpurun_out/f_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-260; done
