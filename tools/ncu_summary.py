"""Extracts the roofline-relevant metrics of every kernel in an .ncu-rep file into a markdown table."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
cols = {
    "Kernel Name": "kernel",
    "launch__grid_size": "grid",
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_%",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_%",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_%active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_%",
    "launch__registers_per_thread": "regs",
}
units = rows[1]
idx = {k: hdr.index(k) for k in cols if k in hdr}
print("| " + " | ".join(cols[k] + (f" [{units[idx[k]]}]" if units[idx[k]] else "") for k in idx) + " |")
print("|" + "---|" * len(idx))
for r in rows[2:]:
    vals = []
    for k in idx:
        v = r[idx[k]]
        if k == "Kernel Name":
            v = "`" + v.replace("void ", "").split("(")[0][:60] + "`"
        else:
            try:
                v = f"{float(v.replace(',', '')):.2f}"
            except ValueError:
                pass
        vals.append(v)
    print("| " + " | ".join(vals) + " |")
