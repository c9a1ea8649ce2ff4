#!/bin/bash
TAG=${1:-r2s}
OUT=gpurun_out
mkdir -p $OUT
echo "== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q --tb=short 2>&1 | tail -25 | tee $OUT/${TAG}_kernel_tests.txt
echo "== model + parity tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -q --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_model_tests.txt
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$name.json").read())
    print(f"$name: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  launches {d['gpu_launches'] // d['steps']}")
except Exception as e:
    print("$name: FAILED", e, open("$OUT/${TAG}_$name.json").read()[-400:])
PY
}
run globbias X=0
run concat ISTNET_GLOBAL_BIAS=0
run globbias2 X=0
echo "== timeline"; timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1; grep -v "chain bwd\|fp bwd" $OUT/${TAG}_timeline.txt | sed -n 8,40p
