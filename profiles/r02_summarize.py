#!/usr/bin/env python
"""profiles/r02_summarize.py — per-launch ncu metrics (profiles/r02_capture.sh, passes 2 and 4) -> committed summaries.

    python profiles/r02_summarize.py gpurun_out/r02_c3_metrics.csv c3 > profiles/r02_c3_metrics.txt
    python profiles/r02_summarize.py gpurun_out/r02_c2_metrics.csv c2 > profiles/r02_c2_metrics.txt

Prints one row per launch (device time, DRAM bytes, issue-slot and L1 data-pipe utilisation, active lanes per instruction, occupancy) and
per-kernel totals over ALL launches of the step, and merges those totals into profiles/r02_traffic.json — the file bench.py reads for
`roofline.traffic` / `roofline.issue` / `roofline.l1` / `roofline.active_lanes` (mean DRAM bytes per launch over the same population of
launches its algorithmic bytes per launch are averaged over; the percentages are duration-weighted means)."""
import collections
import csv
import json
import os
import sys

path, workload = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; kid, kn, mn, mu, mv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
launch = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("b2r::", "").split("<")[0]
    d = launch.setdefault(r[kid], {"kernel": name})
    v = float(r[mv].replace(",", ""))
    u = r[mu]
    if r[mn] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
    if r[mn].startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[mn]] = v
K = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "l1": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "inst": "smsp__inst_executed.sum",
     "occ": "sm__warps_active.avg.pct_of_peak_sustained_active", "fma": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "alu": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "l2hit": "lts__t_sector_hit_rate.pct", "l1hit": "l1tex__t_sector_hit_rate.pct"}
print(f"# {path}: {len(launch)} launches of one {workload.upper()} step (ncu --metrics ..., --clock-control none; profiles/r02_capture.sh)")
print(f"{'kernel':22s} {'us':>9s} {'dram MB':>9s} {'issue %':>8s} {'L1 pipe %':>9s} {'lanes':>6s} {'occ %':>6s} {'fma %':>6s} {'alu %':>6s} {'L1 hit':>6s} {'L2 hit':>6s} {'Minst':>8s}")
tot = collections.OrderedDict()
T = sum(d[K["t"]] for d in launch.values())
for d in launch.values():
    g = lambda k: d.get(K[k], float("nan"))
    print(f"{d['kernel']:22s} {g('t'):9.1f} {(g('rd') + g('wr')) / 1e6:9.2f} {g('issue'):8.1f} {g('l1'):9.1f} {g('lanes'):6.2f} {g('occ'):6.1f} {g('fma'):6.1f} {g('alu'):6.1f} {g('l1hit'):6.1f} {g('l2hit'):6.1f} {g('inst') / 1e6:8.1f}")
    a = tot.setdefault(d["kernel"], collections.defaultdict(float))
    a["n"] += 1; a["t"] += g("t"); a["bytes"] += g("rd") + g("wr"); a["inst"] += g("inst"); a["tinst"] += g("inst") * g("lanes")
    for k in ("issue", "l1", "occ", "fma", "alu"):
        a[k] += g(k) * g("t")
print()
print(f"{'kernel (all launches)':22s} {'n':>3s} {'total us':>9s} {'share':>6s} {'MB/launch':>10s} {'issue %':>8s} {'L1 pipe %':>9s} {'lanes':>6s} {'occ %':>6s} {'fma %':>6s} {'alu %':>6s}")
out = {}
# bench.py books the bounce-0 packet launch with the closest-hit traversal (one KK_CLOSEST launch per bounce) and the k_brute_finish
# launches with the brute-force bounce kernel (KK_BRUTE): the JSON rows "k_intersect_closest" / "k_bounce_brute" are the sums over all
# those launches of the step (the same population bench.py averages its algorithmic bytes over); the parts are listed on their own as well
MERGE = {"k_intersect_closest": ("k_intersect_packet", "k_intersect_closest_per_lane"), "k_bounce_brute": ("k_brute_finish", "k_bounce_brute_per_bounce")}
renamed = {}
for main, (extra, alone) in MERGE.items():
    if main in tot and extra in tot:
        m = collections.defaultdict(float)
        for src in (extra, main):
            for q, v in tot[src].items():
                m[q] += v
        tot[f"{main} (all {int(m['n'])} launches incl. {extra})"] = m
        renamed[main] = alone
for k, a in tot.items():
    w = lambda q: a[q] / a["t"]
    print(f"{k[:22]:22s} {int(a['n']):3d} {a['t']:9.1f} {100 * a['t'] / T:5.1f}% {a['bytes'] / a['n'] / 1e6:10.2f} {w('issue'):8.1f} {w('l1'):9.1f} {a['tinst'] / a['inst']:6.2f} {w('occ'):6.1f} {w('fma'):6.1f} {w('alu'):6.1f}")
    jk = k.split(" (")[0] if " (all " in k else renamed.get(k, k)
    out[jk] = {"launches": int(a["n"]), "dram_bytes_per_launch": a["bytes"] / a["n"], "device_us_per_step": a["t"], "issue_slots_busy_pct": w("issue"), "l1_data_pipe_pct": w("l1"),
              "active_lanes_per_instruction": a["tinst"] / a["inst"], "achieved_occupancy_pct": w("occ"),
              "source": f"profiles/r02_{workload}_metrics.txt: ncu --metrics (dram__bytes_read/write.sum, smsp__issue_active, l1tex__data_pipe_lsu_wavefronts, smsp__thread_inst_executed_per_inst_executed) --clock-control none over ALL {int(a['n'])} launches of the kernel in one {workload.upper()} step; percentages duration-weighted"}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json")
allw = json.load(open(dst)) if os.path.exists(dst) else {}
allw[workload] = out
json.dump(allw, open(dst, "w"), indent=1)
