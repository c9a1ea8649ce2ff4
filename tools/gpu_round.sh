#!/bin/bash
# One GPU session: unit tests of the new kernels first (each under its own timeout), then the full -m gpu suite, a short bench,
# the launch list and the phase timeline.  Everything lands in gpurun_out/.  usage: tools/gpu_round.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== sa_fused tests"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py -q -x --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_t_sa.log
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pointops.py -q --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_t_kernels.log
echo "== model tests"; timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -q --tb=short -s 2>&1 | tail -60 | tee $OUT/${TAG}_t_model.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 30 --warmup 5 2>&1 | tail -5 | tee $OUT/${TAG}_bench.log
echo "== timeline"; timeout 600 python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1; tail -3 $OUT/${TAG}_timeline.txt
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv 60 > $OUT/${TAG}_launches_summary.txt 2>&1; head -30 $OUT/${TAG}_launches_summary.txt
