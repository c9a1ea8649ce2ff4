#!/bin/bash
# Launch list of the CAPTURED step (ncu profiles the kernel nodes of a replayed CUDA graph one by one).  usage: tools/gpu_graph_launches.sh <tag>
TAG=${1:-r2t}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 30000 --csv --log-file $OUT/${TAG}_graph_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
python - <<PY
import csv, re, collections
lines = [l for l in open("$OUT/${TAG}_graph_launches.csv", newline="") if l.startswith('"')]
rows = []
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1000))
pb = [i for i, r in enumerate(rows) if "prep_weight_batch" in r[0]]
print("kernels captured", len(rows), "steps seen", len(pb))
seg = rows[pb[-2]:pb[-1]] if len(pb) >= 2 else rows
agg = collections.defaultdict(lambda: [0, 0.0])
for n, us in seg:
    short = re.sub(r"\(.*", "", n.replace("<unnamed>::", "").replace("(anonymous namespace)::", ""))
    f = re.findall(r"(CUDAFunctor_\w+|\w+Functor|direct_copy_kernel_cuda|MeanOps|NormTwoOps|\w+Ops\b|CatArrayBatchedCopy\w*|FillFunctor)", n)
    key = re.sub(r"<.*", "", short)[:44] + (" [" + f[0] + "]" if f and "at::" in n else "")
    a = agg[key]; a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
with open("$OUT/${TAG}_graph_launches_summary.txt", "w") as fo:
    fo.write(f"one replayed step (last complete one in the capture): {len(seg)} kernel nodes, {tot / 1000:.3f} ms summed (cold caches, serialised)\n")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fo.write(f"{us / 1000:9.3f} ms {100 * us / tot:5.1f}%  x{c:<4d} {k}\n")
print(open("$OUT/${TAG}_graph_launches_summary.txt").read()[:3500])
PY
rm -f $OUT/${TAG}_graph_launches.csv
