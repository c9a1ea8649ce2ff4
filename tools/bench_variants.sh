#!/usr/bin/env bash
# usage: tools/bench_variants.sh "NAME=VAL ..." ["NAME=VAL ..."] ...  — one short bench per environment variant
for v in "$@"; do
  out=$(env $v python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1)
  echo "[$v] $(echo "$out" | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['value'],1),'inst/s', round(d['ms_per_step'],2),'ms')
except Exception as e:
    print('ERR', e)")"
  echo "$out" | grep -v '^{' | tail -3
done
