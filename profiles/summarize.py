#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the small, committed summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_c2.csv  > profiles/r01_c2_launches.txt
    python profiles/summarize.py full gpurun_out/prof_c2.ncu-rep      > profiles/r01_c2_ncu_full.txt

`launches`: per-kernel totals, share of the step and launch counts from the `--metrics gpu__time_duration.sum` pass
(cold-cache, serialised: compare SHARES, not absolutes). `full`: the handful of `--set full` metrics the roofline uses
(DRAM bytes, throughput %, occupancy, issue utilisation, SIMT efficiency, stall samples), per captured launch.
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]; kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt, seq = collections.defaultdict(float), collections.Counter(), []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("b2r::", "")
        v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r[mu], 1.0)
        tot[name] += v; cnt[name] += 1; seq.append((name, v))
    T = sum(tot.values())
    print(f"# {path}: {len(seq)} launches, {T / 1e3:.3f} ms total device time (ncu-serialised)")
    print(f"{'kernel':46s} {'launches':>8s} {'total us':>12s} {'share':>7s} {'avg us':>10s}")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"{k:46s} {cnt[k]:8d} {v:12.1f} {100 * v / T:6.1f}% {v / cnt[k]:10.1f}")


WANT = [
    ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 % of peak"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instr"), ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (unused by design)"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall long_scoreboard"), ("smsp__pcsamp_warps_issue_stalled_wait", "stall wait"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall short_scoreboard"), ("smsp__pcsamp_warps_issue_stalled_barrier", "stall barrier"),
    ("smsp__pcsamp_warps_issue_stalled_branch_resolving", "stall branch"), ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall not_selected"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "issued (selected)"), ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall math_throttle"),
    ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall lg_throttle"), ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall mio_throttle"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    kn = h.index("Kernel Name")
    print(f"# {path}: {len(data)} captured launches (ncu --set full --clock-control none)")
    names = [r[kn].split("(")[0].replace("void ", "").replace("b2r::", "") for r in data]
    print(f"{'metric':36s} {'unit':10s} " + " ".join(f"{n[:24]:>24s}" for n in names))
    for key, label in WANT:
        if key in h:
            i = h.index(key)
            print(f"{label:36s} {units[i][:10]:10s} " + " ".join(f"{r[i][:24]:>24s}" for r in data))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
