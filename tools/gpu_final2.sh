#!/bin/bash
# Lean round-end session: full -m gpu suite, smoke, default bench line (+ kernel table), reference arm, timelines, launch list.
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_gpu_tests.txt
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
echo "== bench"; ISTNET_KERNEL_TABLE=$OUT/${TAG}_kernel_table.txt timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/${TAG}_bench.log | cut -c1-400
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref.log | cut -c1-300
echo "== timeline"; timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1; tail -2 $OUT/${TAG}_timeline.txt
echo "== timeline serial"; ISTNET_STREAMS=0 ISTNET_WGRAD_STREAM=0 ISTNET_SA_FORK=0 timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline_serial.txt 2>&1; tail -2 $OUT/${TAG}_timeline_serial.txt
for c in cfg3 cfg4; do echo "== $c"; timeout 400 python bench.py --config $c --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/${TAG}_$c.json | cut -c1-200; done
echo "== eager"; timeout 400 python bench.py --eager --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_eager.json | cut -c1-200
echo "== launches (eager step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --eager --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv 60 > $OUT/${TAG}_launches_summary.txt 2>&1; head -8 $OUT/${TAG}_launches_summary.txt
rm -f $OUT/${TAG}_launches.csv
