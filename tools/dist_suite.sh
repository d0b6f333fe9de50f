#!/usr/bin/env bash
# tools/dist_suite.sh NGPU TAG : the sharded evaluation on NGPU GPUs of this box, both back-ends, C3 and C4
# (results under gpurun_out/dist_TAG_*.json)
set -u
G=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511"
for wl in c3 c4; do
  for nb in ${NBS:-1024 2048}; do
    timeout 300 python tools/dist_bench.py --workload $wl --ngpu $G --nb $nb --reps 3 > gpurun_out/dist_${TAG}_local_${wl}_${nb}.json 2> gpurun_out/dist_${TAG}_local_${wl}_${nb}.err
    timeout 300 $TR tools/dist_bench.py --workload $wl --backend nccl --nb $nb --reps 3 > gpurun_out/dist_${TAG}_nccl_${wl}_${nb}.json 2> gpurun_out/dist_${TAG}_nccl_${wl}_${nb}.err
  done
done
echo suite done
