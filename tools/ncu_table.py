"""One line per profiled launch of an .ncu-rep (read on the CPU box): kernel, grid, duration, DRAM traffic and achieved GB/s against the
measured HBM peak (MEASURED_PEAKS.json), tensor-pipe %, SM throughput %, issue-slot %, achieved occupancy, registers.
usage: python tools/ncu_table.py file.ncu-rep [--group]   (--group: aggregate per kernel name + template arguments)"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
COLS = OrderedDict([
    ("us", "gpu__time_duration.sum"), ("rd_MB", "dram__bytes_read.sum"), ("wr_MB", "dram__bytes_write.sum"),
    ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), ("sm%", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
    ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
])
# fallbacks for reports captured with --section SpeedOfLight / MemoryWorkloadAnalysis instead of --set full
ALT = {"tensor%": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "issue%": "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
       "GB/s": "dram__bytes.sum.per_second"}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3,
        "byte/s": 1e-9, "Kbyte/s": 1e-6, "Mbyte/s": 1e-3, "Gbyte/s": 1.0, "Tbyte/s": 1e3}


def num(v, unit):
    try:
        return float(v.replace(",", "")) * UNIT.get(unit, 1.0)
    except ValueError:
        return float("nan")


def main():
    f = sys.argv[1]
    group = "--group" in sys.argv
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = {k: hdr.index(m) for k, m in COLS.items() if m in hdr}
    for k, m in ALT.items():
        if k not in idx and m in hdr:
            idx[k] = hdr.index(m)
    recs = []
    for r in rows[2:]:
        name = r[ki].replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
        name = re.sub(r"\(.*", "", name)
        rec = {"name": name}
        for k, i in idx.items():
            rec[k] = num(r[i], units[i])
        if "rd_MB" in rec:
            rec["GB/s"] = (rec.get("rd_MB", 0) + rec.get("wr_MB", 0)) / max(rec.get("us", 1e-9), 1e-9) * 1e3
        else:  # only the rate was collected: traffic = rate x duration
            rec["rd_MB"], rec["wr_MB"] = rec.get("GB/s", 0.0) * rec.get("us", 0.0) * 1e-3, 0.0
        rec["hbm_frac"] = rec["GB/s"] / HBM
        recs.append(rec)
    keys = ["us", "rd_MB", "wr_MB", "GB/s", "hbm_frac", "tensor%", "sm%", "issue%", "l2hit%", "warps%", "regs", "grid"]
    print(f"# {os.path.basename(f)}: {len(recs)} launches; GB/s = (dram read + write) / duration; hbm_frac against the measured {HBM:.0f} GB/s")
    print(f"{'kernel':58s} " + " ".join(f"{k:>8s}" for k in keys) + ("      n" if group else ""))
    if group:
        agg = OrderedDict()
        for r in recs:
            agg.setdefault(r["name"], []).append(r)
        for name, rs in sorted(agg.items(), key=lambda kv: -sum(x.get("us", 0) for x in kv[1])):
            tot = sum(x.get("us", 0) for x in rs)
            line = {k: sum(x.get(k, 0) for x in rs) / len(rs) for k in keys}
            line["us"], line["rd_MB"], line["wr_MB"] = tot, sum(x.get("rd_MB", 0) for x in rs), sum(x.get("wr_MB", 0) for x in rs)
            line["GB/s"] = (line["rd_MB"] + line["wr_MB"]) / max(tot, 1e-9) * 1e3
            line["hbm_frac"] = line["GB/s"] / HBM
            print(f"{name[:58]:58s} " + " ".join(f"{line[k]:8.2f}" if k in ("hbm_frac",) else f"{line[k]:8.1f}" for k in keys) + f" {len(rs):6d}")
    else:
        for r in recs:
            print(f"{r['name'][:58]:58s} " + " ".join(f"{r.get(k, float('nan')):8.2f}" if k in ("hbm_frac",) else f"{r.get(k, float('nan')):8.1f}" for k in keys))


main()
