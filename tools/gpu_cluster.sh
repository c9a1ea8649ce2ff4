#!/bin/bash
# Bring-up of the CTA-pair (cluster 2, weight-tile multicast) variant of conv_gemm_tc_kernel.  Everything under hard timeouts.
TAG=${1:-r2x}
OUT=gpurun_out
mkdir -p $OUT
echo "== conv tests, ISTNET_CG_CLUSTER=2"
ISTNET_CG_CLUSTER=2 timeout -s KILL 150 python -m pytest tests/test_gpu_kernels.py -q --tb=short -x -k "conv_gemm or unit or basic_block or global_feature" 2>&1 | tail -12 | tee $OUT/${TAG}_tests.txt
rc=${PIPESTATUS[0]}
echo "pytest rc=$rc"
if [ "$rc" != "0" ]; then nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv; exit 0; fi
echo "== model tests, ISTNET_CG_CLUSTER=2"
ISTNET_CG_CLUSTER=2 timeout -s KILL 240 python -m pytest tests/test_gpu_model.py -q --tb=short -x 2>&1 | tail -6 | tee -a $OUT/${TAG}_tests.txt
run() {
  local name=$1; shift
  env "$@" timeout -s KILL 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > $OUT/${TAG}_$name.json
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$name.json").read())
    print(f"$name: {d['value']:.1f} inst/s  {d['ms_per_step']:.3f} ms/step  roofline {d['roofline']['executed_mma_tflops']:.0f} MMA-TF/s")
except Exception as e:
    print("$name: FAILED", e, open("$OUT/${TAG}_$name.json").read()[-300:])
PY
}
run cluster2 ISTNET_CG_CLUSTER=2 ISTNET_KERNEL_TABLE=$OUT/${TAG}_table_cluster2.txt
run single X=0
grep "cin=512 cout=512 k=1\|cin=1024 cout=256 k=3\|24x24 cin=512 cout=512 k=3" $OUT/${TAG}_table_cluster2.txt | grep conv_gemm | sort -k1 -g | head -12
