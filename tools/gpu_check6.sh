#!/bin/bash
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
echo "== kernel + model tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -q --tb=short 2>&1 | tail -6 | tee $OUT/${TAG}_tests.txt
for i in 1 2; do
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_b$i.json
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_b$i.json").read())
print(f"run $i: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}")
PY
done
