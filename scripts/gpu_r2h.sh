#!/bin/bash
# round 2, eighth GPU pass (2 GPUs): sharded path through the library's communicator, bench at N = 2; GELU timing on GPU 0
set -u
O=gpurun_out
mkdir -p $O
run() { local name=$1 t=$2; shift 2; timeout $t "$@" > $O/r2h_$name.log 2> $O/r2h_$name.err; echo "$name exit $?" >> $O/r2h_summary.txt; }
: > $O/r2h_summary.txt
nvidia-smi --query-gpu=index,name --format=csv > $O/r2h_gpus.txt 2>&1
run gemm5 600 python tests/gpu_selftest.py gemm --impl 5
run pytest_multi 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu
run bench_n2 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3
ZETT_GATHER=nccl run bench_n2_nccl 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 3 --warmup 3
run bench_ref_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1
ZETT_GEMM_PROF=1 run sweep 600 python tests/gpu_selftest.py sweep --sweep-terms 2 --mnk "53248,8192,4096;54000,1536,768"
run fwd 900 python tests/gpu_selftest.py forward --impl 0 --configs tiny,tiny_lang,xlmr,mistral
