#!/bin/bash
# Leanest round-end check of the committed state: full -m gpu suite, smoke, default bench line.
TAG=${1:-lean}
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -6 | tee $OUT/${TAG}_gpu_tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/${TAG}_bench.log | cut -c1-330
