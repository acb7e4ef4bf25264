#!/bin/bash
# Wide-tile heuristic (K >= 4096, whole waves): benches of the three BASELINE shapes, then the driver's GPU test command.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
rm -f gpurun_out/summary.txt
timeout 400 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tinyllama_${R}.json 2> gpurun_out/bench_tinyllama.err
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mistral_${R}.json 2> gpurun_out/bench_mistral.err
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 300 python bench.py --config xlmr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_xlmr_${R}.json 2> gpurun_out/bench_xlmr.err
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
timeout 400 python -m pytest tests/ -x -q -m gpu > gpurun_out/i8_pytest_gpu.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/i8_pytest_gpu.log
for f in gpurun_out/bench_*_${R}.json; do echo $f; tail -n 1 $f | cut -c1-200; done
