#!/bin/bash
# 2-GPU sanity of the driver's launch line (own arm + reference arm), each under its own timeout.
TAG=${1:-r2n2}
OUT=gpurun_out
mkdir -p $OUT
N=${2:-2}
echo "== own arm N=$N"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | tail -4 | tee $OUT/${TAG}_bench.log | cut -c1-500
echo "== reference arm N=$N"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -2 | tee $OUT/${TAG}_bench_ref.log | cut -c1-300
