#!/bin/bash
# round 2, eighteenth GPU pass (8 GPUs): final build at N = 8 -- sharded == single GPU test, bench with both transports
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2r_$name.log 2> $O/r2r_$name.err; echo "$name exit $?" >> $O/r2r_summary.txt; }
: > $O/r2r_summary.txt
run pytest_multi 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu
run bench_n8 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3
ZETT_GATHER=nccl run bench_n8_nccl 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 --no-extra
