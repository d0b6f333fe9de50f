#!/usr/bin/env bash
set -u
G=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
run_nccl() { timeout 300 $TR tools/dist_bench.py --workload $1 --backend nccl --nb $2 --reps 4 > gpurun_out/dist_${TAG}_nccl_$1_$2.json 2> gpurun_out/dist_${TAG}_nccl_$1_$2.err; }
run_nccl c3 1024; run_nccl c4 2048; run_nccl c3 512; run_nccl c4 1024
echo suite done
