#!/bin/bash
# Transposed (whole-line) epilogue stores: GPU tests, sustained probes with / without stores, benches.
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/i6_pytest_gpu.log 2>&1
echo "pytest -m gpu exit $?" >> gpurun_out/summary.txt
ZETT_SUSTAINED_ONLY="f16+2xe5m2,bf16x3 256,bf16 single" timeout 600 python tests/gpu_selftest.py sustained --mnk "53248,12288,4096;65536,2304,768" > gpurun_out/i6_sustained.log 2>&1
echo "sustained exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/i6_bench_mistral.log 2>&1
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 4 --warmup 3 --gemm-impl 2 --no-cpu-baseline > gpurun_out/i6_bench_mistral_i2.log 2>&1
echo "bench mistral impl 2 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config xlmr --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/i6_bench_xlmr.log 2>&1
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config tinyllama --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/i6_bench_tinyllama.log 2>&1
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/i6_pytest_gpu.log
grep -h '"kind": "sustained"' gpurun_out/i6_sustained.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('%6d %5d %-55s ms %7.3f  MHz %6.0f  W %5.0f' % (d['m'], d['n'], d.get('label',d['impl']), d['ms'], d.get('sm_mhz',0), d.get('power_w',0)))
"
for f in gpurun_out/i6_bench_*.log; do echo $f; tail -n 1 $f | cut -c1-200; done
