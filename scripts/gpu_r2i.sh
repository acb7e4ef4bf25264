#!/bin/bash
# round 2, ninth GPU pass (8 GPUs): bench at N = 8 with both gather transports (incl. the 256k-row workload in `extra`)
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2i_$name.log 2> $O/r2i_$name.err; echo "$name exit $?" >> $O/r2i_summary.txt; }
: > $O/r2i_summary.txt
nvidia-smi --query-gpu=index,name --format=csv > $O/r2i_gpus.txt 2>&1
run bench_n8 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3
ZETT_GATHER=nccl run bench_n8_nccl 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 3 --warmup 3 --no-extra
