#!/bin/bash
# GPU tests + bench + ncu launch list + one full capture of the GEMM kernel. Logs under gpurun_out/.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mistral.log 2>&1
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config xlmr --steps 5 --warmup 3 > gpurun_out/bench_xlmr.log 2>&1
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tinyllama.log 2>&1
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 130 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 100 -c 3 -f -o gpurun_out/gemm_${R} \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none -k regex:gather_rescale -s 3 -c 1 -f -o gpurun_out/gather_${R} \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gather.log 2>&1
echo "ncu gather exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 5 gpurun_out/pytest_gpu.log
tail -n 1 gpurun_out/bench_mistral.log
