#!/bin/bash
# Warp-per-row LayerNorm + float4 attention lanes (row-wise epilogue stores restored): GPU tests, launch lists of one
# pass (XLM-R and Mistral shapes), benches.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/i7_pytest_gpu.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 140 --csv --log-file gpurun_out/launches_xlmr_${R}.csv \
  python bench.py --config xlmr --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/i7_ncu_xlmr.log 2>&1
echo "ncu xlmr launches exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 130 --csv --log-file gpurun_out/launches_${R}.csv \
  python bench.py --rows 16384 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/i7_ncu_mistral.log 2>&1
echo "ncu mistral launches exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mistral_${R}.json 2> gpurun_out/bench_mistral.err
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tinyllama_${R}.json 2> gpurun_out/bench_tinyllama.err
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config xlmr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_xlmr_${R}.json 2> gpurun_out/bench_xlmr.err
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/i7_pytest_gpu.log
for f in gpurun_out/bench_*_${R}.json; do echo $f; tail -n 1 $f | cut -c1-200; done
