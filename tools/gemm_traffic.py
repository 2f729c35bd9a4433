#!/usr/bin/env python
"""DRAM traffic of the GEMM launches of one step from an ncu CSV (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum,
-k regex:gemm_bf16, tools/profile_step.py):  python tools/gemm_traffic.py gemm_dram.csv profiles/r1_gemm_dram.json
bench.py reads the JSON for roofline.traffic (bytes per launch, averaged over the same launches `achieved` averages over)."""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ik, im, iv, iu, iid = (h.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
per = {}
for r in rows[1:]:
    per.setdefault(r[iid], {"kernel": r[ik]})[r[im]] = float(r[iv].replace(",", "")) * mult.get(r[iu], 1.0)
n = len(per)
rd = sum(d.get("dram__bytes_read.sum", 0.0) for d in per.values())
wr = sum(d.get("dram__bytes_write.sum", 0.0) for d in per.values())
ns = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
by = {}
for d in per.values():
    name = d["kernel"].split("gemm_bf16_kernel")[1].split("(")[0] if "gemm_bf16_kernel" in d["kernel"] else d["kernel"]
    e = by.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    e["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_bf16 python tools/profile_step.py (one data2vec step, B=128)",
       "launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "bytes_per_launch": (rd + wr) / max(n, 1), "kernel_ms": ns / 1e6,
       "by_instantiation": by}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps({k: out[k] for k in ("launches", "dram_read_bytes", "dram_write_bytes", "bytes_per_launch", "kernel_ms")}))
