#!/bin/bash
# First-contact GPU run: every kernel variant in its own process, logs under gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))" > gpurun_out/torch.txt 2>&1
for impl in 3 1 2; do
  timeout 600 python tests/gpu_selftest.py gemm --impl $impl > gpurun_out/gemm_impl$impl.log 2>&1
  echo "gemm impl $impl exit $?" >> gpurun_out/summary.txt
done
for impl in 3 1 2; do
  timeout 900 python tests/gpu_selftest.py forward --impl $impl > gpurun_out/fwd_impl$impl.log 2>&1
  echo "forward impl $impl exit $?" >> gpurun_out/summary.txt
done
timeout 900 python tests/gpu_selftest.py forward --impl 3 --configs xlmr > gpurun_out/fwd_xlmr_impl3.log 2>&1
echo "forward xlmr impl 3 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 1 --configs xlmr,tinyllama,mistral > gpurun_out/fwd_big_impl1.log 2>&1
echo "forward big impl 1 exit $?" >> gpurun_out/summary.txt
timeout 900 python tests/gpu_selftest.py forward --impl 2 --configs xlmr,tinyllama,mistral > gpurun_out/fwd_big_impl2.log 2>&1
echo "forward big impl 2 exit $?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --config xlmr --steps 3 --warmup 3 --gemm-impl 1 --no-cpu-baseline > gpurun_out/bench_xlmr_impl1.log 2>&1
echo "bench xlmr impl1 exit $?" >> gpurun_out/summary.txt
timeout 1200 python bench.py --steps 3 --warmup 3 --gemm-impl 1 --no-cpu-baseline > gpurun_out/bench_mistral_impl1.log 2>&1
echo "bench mistral impl1 exit $?" >> gpurun_out/summary.txt
timeout 1200 python bench.py --steps 3 --warmup 3 --gemm-impl 2 --no-cpu-baseline > gpurun_out/bench_mistral_impl2.log 2>&1
echo "bench mistral impl2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 3 gpurun_out/gemm_impl*.log gpurun_out/fwd_impl*.log
