#!/bin/bash
# Short follow-up session: the tests that changed, their printed error tables, A/B of the SA scale fork, a fresh default bench line.
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity (full shapes, printed tables)"; timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -s --tb=short 2>&1 | grep -v "^$" | tail -30 | tee $OUT/${TAG}_parity_full.txt
echo "== cfg4 parity with 3-plane backward (evidence for the 2-plane median)"; ISTNET_NSPLIT_BWD=3 timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -s --tb=line -k cfg4 2>&1 | grep -v "^$" | tail -6 | tee $OUT/${TAG}_cfg4_bwd3.txt
echo "== solver / adam / model tests"; timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_sa_fused.py -q -s --tb=short 2>&1 | grep -v "^$" | tail -25 | tee $OUT/${TAG}_model_tests.txt
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
run sa_fork X=0
run sa_nofork ISTNET_SA_FORK=0
run sa_fork_again X=0
echo "== bench (default, full line)"; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/${TAG}_bench.log | cut -c1-300
echo "== timeline"; timeout 300 python tools/timeline.py > $OUT/${TAG}_timeline.txt 2>&1; tail -2 $OUT/${TAG}_timeline.txt
