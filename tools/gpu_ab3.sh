#!/bin/bash
# A/B of the stream-placement switches at the round's final balance.  usage: tools/gpu_ab3.sh <tag>
TAG=${1:-ab3}
OUT=gpurun_out
mkdir -p $OUT
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$name.json").read())
    print(f"$name: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step")
except Exception as e:
    print("$name: FAILED", e, open("$OUT/${TAG}_$name.json").read()[-300:])
PY
}
run default X=0
run main_lo ISTNET_MAIN_HI=0
run heads_main_last ISTNET_HEADS_MAIN_FIRST=0
run wgrad_same_stream ISTNET_WGRAD_STREAM=0
