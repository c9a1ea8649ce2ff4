#!/bin/bash
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
echo "== conv tests (default) + cluster 2"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q --tb=short -k "conv_gemm or unit or global_feature" 2>&1 | tail -3
ISTNET_CG_CLUSTER=2 timeout -s KILL 150 python -m pytest tests/test_gpu_kernels.py -q --tb=short -k "conv_gemm or unit or global_feature" 2>&1 | tail -3
for i in 1 2; do
  ISTNET_KERNEL_TABLE=$OUT/${TAG}_table$i.txt timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_b$i.json
  python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_b$i.json").read())
print(f"run $i: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  conv avg {d['roofline']['avg_launch_ms']*1000:.1f} us  MMA {d['roofline']['executed_mma_tflops']:.0f} TF/s")
PY
done
ISTNET_CG_CLUSTER=2 timeout -s KILL 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_cl2.json
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_cl2.json").read())
print(f"cluster2: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  conv avg {d['roofline']['avg_launch_ms']*1000:.1f} us  MMA {d['roofline']['executed_mma_tflops']:.0f} TF/s")
PY
