#!/bin/bash
# A/B runs of bench.py (cfg1, 20 steps) under single environment switches + the other BASELINE.json configs.  usage: tools/gpu_ab.sh <tag>
TAG=${1:-ab}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, env...
  local name=$1; shift
  echo "== $name"
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$name.json").read())
    print(f"$name: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  launches/step {d.get('gpu_launches')}")
except Exception as e:
    print("$name: FAILED", e, open("$OUT/${TAG}_$name.json").read()[-300:])
PY
}
run default X=0
run prio_none ISTNET_PRIO=none
run prio_points ISTNET_PRIO=points
run fuse_bn_bwd ISTNET_FUSE_BN_BWD=1
run sa_unfused ISTNET_SA_FUSED=0
run no_weight_bank ISTNET_WEIGHT_BANK=0
run torch_heads ISTNET_FUSED_HEADS=0
echo "== cfg3"; timeout 400 python bench.py --config cfg3 --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/${TAG}_cfg3.json | cut -c1-300
echo "== cfg4"; timeout 400 python bench.py --config cfg4 --steps 20 --warmup 5 2>&1 | tail -1 | tee $OUT/${TAG}_cfg4.json | cut -c1-300
echo "== eager"; timeout 400 python bench.py --eager --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_eager.json | cut -c1-300
