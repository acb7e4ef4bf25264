#!/bin/bash
# Round validation on one B200: compute-sanitizer on the tiny configurations (default path and the cluster-of-four
# GEMM), benches of the three BASELINE shapes with the CPU baseline and the torch-eager arm, the reference arm.
set +e
mkdir -p gpurun_out
R=${ROUND_TAG:-r1}
rm -f gpurun_out/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang > gpurun_out/sanitizer_memcheck_${R}.log 2>&1
echo "memcheck default exit $?" >> gpurun_out/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/gpu_selftest.py forward --impl 4 --configs tiny,tiny_lang > gpurun_out/sanitizer_memcheck_impl4_${R}.log 2>&1
echo "memcheck impl 4 exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_mistral_${R}.json 2> gpurun_out/bench_mistral.err
echo "bench mistral exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config tinyllama --steps 5 --warmup 3 > gpurun_out/bench_tinyllama_${R}.json 2> gpurun_out/bench_tinyllama.err
echo "bench tinyllama exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --config xlmr --steps 5 --warmup 3 > gpurun_out/bench_xlmr_${R}.json 2> gpurun_out/bench_xlmr.err
echo "bench xlmr exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --split-terms 3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mistral_bf16x3_${R}.json 2> gpurun_out/bench_mistral_bf16x3.err
echo "bench mistral bf16x3 exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_${R}.json 2> gpurun_out/bench_reference.err
echo "bench reference exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
for f in gpurun_out/bench_*_${R}.json; do echo $f; tail -n 1 $f | cut -c1-220; done
tail -n 2 gpurun_out/sanitizer_memcheck_${R}.log gpurun_out/sanitizer_memcheck_impl4_${R}.log
