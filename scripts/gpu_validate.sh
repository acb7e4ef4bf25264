#!/bin/bash
# full validation: GPU pytest, sanitizer on the tiny configs, benches with baselines
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest gpu exit $?" >> gpurun_out/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_selftest.py forward --impl 2 --configs tiny,tiny_lang > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/summary.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tests/gpu_selftest.py forward --impl 2 --configs tiny > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mistral.log 2>&1
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config xlmr --steps 5 --warmup 3 > gpurun_out/bench_xlmr.log 2>&1
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config tinyllama --steps 5 --warmup 3 > gpurun_out/bench_tinyllama.log 2>&1
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
echo "bench reference exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv --log-file gpurun_out/launches_r1b.csv \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
