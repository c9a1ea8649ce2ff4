#!/bin/bash
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
echo "== adam test"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q --tb=short -k adam 2>&1 | tail -8 | tee $OUT/${TAG}_adam.txt
echo "== cfg4 parity, 3 operand planes everywhere (forward + backward)"; ISTNET_NSPLIT_MAP=none=3 ISTNET_NSPLIT_HEADS=3 ISTNET_NSPLIT_BWD=3 timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -s --tb=line -k cfg4 2>&1 | grep -v "^$" | tail -6 | tee $OUT/${TAG}_cfg4_all3.txt
echo "== cfg4 parity, unfused SA levels"; ISTNET_SA_FUSED=0 timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -s --tb=line -k cfg4 2>&1 | grep -v "^$" | tail -6 | tee $OUT/${TAG}_cfg4_unfused.txt
echo "== cfg4 parity, other seeds"; timeout 600 python - <<'PY' 2>&1 | grep -v "^$" | tail -8 | tee $OUT/${TAG}_cfg4_seeds.txt
import sys
sys.path.insert(0, "tests")
import test_gpu_parity_full as T
for seed in (55, 56, 57):
    try:
        T._check("ist_net", 2, 4096, 192, seed=seed, median_factor=100.0)
    except AssertionError as e:
        print("seed", seed, "assert:", str(e)[:200])
PY
