#!/bin/bash
# Short GPU session: fused-SA tests, model tests, bench, launch list.  usage: tools/gpu_quick.sh <tag>
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== sa_fused tests"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py -q -x --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_t_sa.log
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -q --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_t_model.log
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -3 | cut -c1-400 | tee $OUT/${TAG}_bench.log
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv 30 > $OUT/${TAG}_launches_summary.txt 2>&1; head -24 $OUT/${TAG}_launches_summary.txt
