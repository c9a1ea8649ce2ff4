#!/bin/bash
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
echo "== wgrad + conv unit tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q --tb=short -k "wgrad or unit or basic_block or stem" 2>&1 | tail -15 | tee $OUT/${TAG}_wgrad_tests.txt
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
run pair ISTNET_KERNEL_TABLE=$OUT/${TAG}_kernel_table.txt
run nopair ISTNET_WG_NOPAIR=1
run pair2 X=0
grep "wgrad" $OUT/${TAG}_kernel_table.txt | sort -k1 -g -r | head -12
echo "== model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -q --tb=short 2>&1 | tail -6 | tee $OUT/${TAG}_model_tests.txt
