#!/usr/bin/env python3
r"""profiles/traffic_<commit>.json from an `ncu --set full` (or --metrics dram__bytes_*) CSV of ONE compress of the benchmark
workload: the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the forward-transform and quantise kernels, the
scope bench.py's `roofline` is quoted on.  bench.py prints the figure only while the kernel sources it was captured from
(kernels_sha) are unchanged.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        -k regex:fwd_level_fused\|quantise_kernel --csv --log-file gpurun_out/traffic.csv python tools/profile_encode.py 512 f32 1e-4
    python tools/ncu_traffic.py gpurun_out/traffic.csv 512 3 [encodes in the capture]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path, edge, nlay = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
nenc = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if len(r) > 5 and r[0] == "ID")
k = {n: hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit")}
per = {}
unit_scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
for r in rows:
    if len(r) != len(hdr) or r[0] == "ID":
        continue
    name = r[k["Kernel Name"]]
    kind = "forward" if "fwd_level_fused" in name else ("quantise" if "quantise_kernel" in name else None)
    if kind is None:
        continue
    m, v, u = r[k["Metric Name"]], float(r[k["Metric Value"]].replace(",", "")), r[k["Metric Unit"]]
    d = per.setdefault(kind, {"dram_read": 0.0, "dram_write": 0.0, "time_ns": 0.0, "launches": 0})
    if m == "dram__bytes_read.sum":
        d["dram_read"] += v * unit_scale.get(u, 1)
    elif m == "dram__bytes_write.sum":
        d["dram_write"] += v * unit_scale.get(u, 1)
    elif m == "gpu__time_duration.sum":
        d["time_ns"] += v * unit_scale.get(u, 1)
        d["launches"] += 1
for d in per.values():
    for key in ("dram_read", "dram_write", "time_ns"):
        d[key] /= nenc
    d["launches"] //= nenc
total = sum(d["dram_read"] + d["dram_write"] for d in per.values())
commit = subprocess.run(["git", "rev-parse", "--short=12", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
out = {"edge": edge, "nlay": nlay, "dram_bytes": int(total), "per_kernel_family": per, "commit": commit,
       "kernels_sha": bench.kernels_sha(), "algorithmic_bytes": edge ** 3 * (4 + 8 + 9 * nlay),
       "source": os.path.basename(path), "note": "per compress; ncu serialises and flushes caches between kernels"}
dst = os.path.join(ROOT, "profiles", "traffic_%s.json" % commit)
json.dump(out, open(dst, "w"), indent=1)
print(dst, json.dumps(out)[:400])
