#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for N in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_mistral_n$N.log 2>&1
echo "bench n=$N exit $?" >> gpurun_out/summary.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_n8.log 2>&1
echo "multi check n=8 exit $?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_reference_n8.log 2>&1
echo "reference n=8 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 1 gpurun_out/bench_mistral_n8.log | cut -c 1-400
tail -n 3 gpurun_out/multi_check_n8.log
