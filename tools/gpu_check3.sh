#!/bin/bash
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
echo "== dataprep tests"; timeout 600 python -m pytest tests/test_gpu_dataprep.py -q --tb=short 2>&1 | tail -25 | tee $OUT/${TAG}_dataprep.txt
echo "== dataprep throughput"; timeout 300 python tools/bench_dataprep.py 2>&1 | tail -6 | tee $OUT/${TAG}_dataprep_bench.txt
