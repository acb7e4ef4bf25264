#!/bin/bash
set +e
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
N=${NGPUS:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_mistral_n$N.log 2>&1
echo "bench n=$N exit $?" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --config xlmr > gpurun_out/bench_xlmr_n$N.log 2>&1
echo "bench xlmr n=$N exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "multi check n=$N exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 2 gpurun_out/bench_mistral_n$N.log
tail -n 5 gpurun_out/multi_check_n$N.log
