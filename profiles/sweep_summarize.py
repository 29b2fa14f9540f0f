#!/usr/bin/env python
"""profiles/r02_sweep_build_launches.txt from the ncu CSV of profiles/sweep_build_workload.py (gpu__time_duration.sum per launch): one table per
b2r_upload_scene call (a call starts at its k_morton_keys launch)."""
import collections, csv, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_sweep_build_launches.csv"
rows = []
with open(src) as f:
    for row in csv.DictReader([l for l in f if not l.startswith("==")]):
        rows.append((row["Kernel Name"], float(row["Metric Value"].replace(",", ""))))
starts = [i for i, x in enumerate(rows) if "k_morton_keys" in x[0]]
out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python profiles/sweep_build_workload.py (C3's scene, 100k spheres): per b2r_upload_scene call,",
       "launches and summed device time per kernel (cold-cache, serialised: shares, not absolutes). Calls 0-3: packed tree (B2R_FLAG_GPU_TREE; call 0 = b2r_create's",
       "upload + origin-box refit), calls 4-7: curve sweep tree (| B2R_FLAG_GPU_SAH), calls 8-11: three-axis sweep tree (| B2R_FLAG_GPU_SAH3).", ""]
for k, a in enumerate(starts):
    b = starts[k + 1] if k + 1 < len(starts) else len(rows)
    agg = collections.OrderedDict()
    for name, v in rows[a:b]:
        short = re.sub(r"<.*", "", name.split("(")[0]).split("::")[-1]
        if "Scan" in name: short = "cub::DeviceScan (exclusive sums: run counts, partition places): " + short
        if "Radix" in name: short = "cub::DeviceRadixSort: " + short
        d = agg.setdefault(short, [0, 0.0]); d[0] += 1; d[1] += v
    out.append(f"call {k}")
    for name, (c, v) in agg.items():
        out.append(f"  {name:78s} {c:4d} launches {v / 1000:9.1f} us")
    out.append(f"  {'total':78s} {sum(v[0] for v in agg.values()):4d} launches {sum(v[1] for v in agg.values()) / 1000:9.1f} us")
print("\n".join(out))
