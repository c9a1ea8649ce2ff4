#!/bin/bash
# A/B of the fused-SA grid sizes (CTAs per SM) on the whole step.  usage: tools/gpu_ab2.sh <tag>
TAG=${1:-ab2}
OUT=gpurun_out
mkdir -p $OUT
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$name.json").read())
    print(f"$name: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}")
except Exception as e:
    print("$name: FAILED", e, open("$OUT/${TAG}_$name.json").read()[-300:])
PY
}
run default X=0
run f1b1 ISTNET_SA_FWD_CTAS=1 ISTNET_SA_BWD_CTAS=1
run f1b2 ISTNET_SA_FWD_CTAS=1 ISTNET_SA_BWD_CTAS=2
run f2b2 ISTNET_SA_FWD_CTAS=2 ISTNET_SA_BWD_CTAS=2
run f2b4 ISTNET_SA_FWD_CTAS=2 ISTNET_SA_BWD_CTAS=4
run f4b4 ISTNET_SA_FWD_CTAS=4 ISTNET_SA_BWD_CTAS=4
run default2 X=0
echo "== sa tests at 1 / 4 CTAs per SM"; ISTNET_SA_FWD_CTAS=1 ISTNET_SA_BWD_CTAS=1 timeout 300 python -m pytest tests/test_gpu_sa_fused.py -q --tb=line 2>&1 | tail -3
ISTNET_SA_FWD_CTAS=4 ISTNET_SA_BWD_CTAS=4 timeout 300 python -m pytest tests/test_gpu_sa_fused.py -q --tb=line 2>&1 | tail -3
