"""Dumps the judged metrics of an .ncu-rep (read on the CPU box) to text: usage ncu_summary.py file.ncu-rep [more...]"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"== {f}")
    ki = hdr.index("Kernel Name")
    print("kernels:", sorted(set(r[ki].split("(")[0] for r in rows[2:])))
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"{h} [{units[i]}]: {[r[i] for r in rows[2:]]}")
