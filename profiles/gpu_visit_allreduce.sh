#!/bin/bash
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for cfg in "" "NCCL_ALGO=NVLS" "NCCL_ALGO=NVLSTree" "NCCL_ALGO=Ring" "NCCL_ALGO=Tree" "NCCL_ALGO=Ring NCCL_PROTO=Simple"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29711 profiles/allreduce_bench.py 2>&1 | grep -E "^world|rror" | head -6
done | tee gpurun_out/allreduce_bench.txt
