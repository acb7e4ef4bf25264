#!/bin/bash
# round 2 (2 GPUs): soft-failing peer registration -- sharded == single GPU test and a short bench at N = 2
set -u
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > $O/r2v_pytest_multi.log 2>&1; echo "pytest exit $?" > $O/r2v_summary.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra > $O/r2v_bench_n2.log 2> $O/r2v_bench_n2.err; echo "bench exit $?" >> $O/r2v_summary.txt
