#!/bin/bash
# A/B of the weight-gradient kernel's tuning knobs after the tap-pairing change.  usage: tools/gpu_ab4.sh <tag>
TAG=${1:-ab4}
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
run wg_kscap74 ISTNET_WG_KSCAP=74
run wg_smem110 ISTNET_WG_SMEM_KB=110
run wg_pix32 ISTNET_WG_PIX=32
