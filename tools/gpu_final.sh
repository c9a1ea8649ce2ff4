#!/bin/bash
# Round-end validation session: full -m gpu suite, smoke, the default bench line, phase timelines (overlapped / serial), the A/B and
# cfg3 / cfg4 / eager lines (tools/gpu_ab.sh), inference latencies and the launch lists.  usage: tools/gpu_final.sh <tag>
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee $OUT/${TAG}_gpu_tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -5 | tee $OUT/${TAG}_bench.log | cut -c1-600
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee $OUT/${TAG}_bench_ref.log | cut -c1-400
echo "== timeline"; timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1; tail -2 $OUT/${TAG}_timeline.txt
echo "== timeline serial"; ISTNET_STREAMS=0 ISTNET_WGRAD_STREAM=0 ISTNET_SA_FORK=0 timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline_serial.txt 2>&1; tail -2 $OUT/${TAG}_timeline_serial.txt
bash tools/gpu_ab.sh ${TAG}
echo "== inference latency"; timeout 300 python tools/bench_infer.py 2>&1 | tail -14 | tee $OUT/${TAG}_infer.txt
echo "== launches (eager step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv 60 > $OUT/${TAG}_launches_summary.txt 2>&1; head -12 $OUT/${TAG}_launches_summary.txt
du -sh $OUT
