"""From an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of one
training step (tools/profile_step.py): DRAM traffic of the conv kernels, total and average per launch ->
profiles/r01_conv_traffic.json (read by bench.py for roofline.traffic).  Usage: conv_traffic.py launches.csv out.json"""
import csv
import json
import re
import sys
from collections import defaultdict

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
per = defaultdict(lambda: defaultdict(float))  # launch id -> metric -> value
name = {}
for r in csv.DictReader(lines):
    m = r["Metric Name"]
    if m not in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if m.startswith("dram"):
        v *= SCALE.get(r["Metric Unit"], 1.0)
    per[r["ID"]][m] = v
    name[r["ID"]] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
conv = [i for i, n in name.items() if "conv_fprop_kernel" in n or "conv_wgrad_kernel" in n]
rd = sum(per[i]["dram__bytes_read.sum"] for i in conv)
wr = sum(per[i]["dram__bytes_write.sum"] for i in conv)
allrd = sum(v["dram__bytes_read.sum"] for v in per.values())
allwr = sum(v["dram__bytes_write.sum"] for v in per.values())
out = {
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (cold L2 per kernel), tools/profile_step.py 16 1",
    "conv_launches": len(conv),
    "conv_dram_bytes_per_step": rd + wr,
    "conv_dram_bytes_per_launch_avg": (rd + wr) / max(1, len(conv)),
    "conv_dram_read_bytes_per_step": rd,
    "conv_dram_write_bytes_per_step": wr,
    "all_kernels_dram_bytes_per_step": allrd + allwr,
    "all_kernels_launches": len(per),
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
