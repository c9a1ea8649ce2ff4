#!/bin/bash
# Diagnostic: the captured-NCCL variant (ISTNET_GRAPH_NCCL=1) at N=2 with stage markers, hard 110 s limit.
OUT=gpurun_out; mkdir -p $OUT
ISTNET_GRAPH_NCCL=1 ISTNET_TRACE_STAGES=1 NCCL_DEBUG=WARN timeout -s KILL 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2o_graphnccl.log 2>&1
echo "exit $?"; grep -v "^$" $OUT/r2o_graphnccl.log | tail -25 | cut -c1-400
nvidia-smi --query-gpu=index,utilization.gpu,memory.used --format=csv
