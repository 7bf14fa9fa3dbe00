#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.

  python profiles/summarize.py <tag>      # expects gpurun_out/<tag>_launches.csv and <tag>_<kernel>.ncu-rep
"""
import csv
import collections
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
out = open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w")


def w(s=""):
    print(s)
    out.write(s + "\n")


# ---- launch list
path = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)   # -> us
        name = r[ik].split("(")[0].replace("void ", "")[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    w(f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)")
    w("command: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-refcuda --cameras 1   (first 400 launches; every launch is camera 0)")
    w()
    w("| kernel | launches | total us | share |")
    w("|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    w()

# ---- launch list of the train step (optional)
path = os.path.join(G, f"{tag}train_launches.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)
        name = r[ik].split("(")[0].replace("void ", "")[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    w(f"# {tag}: ncu launch list of the TRAIN STEP (bench.py --workload train_step --steps 2 --warmup 3 --no-e2e --cameras 1 --refine-every 3)")
    w()
    w("| kernel | launches | total us | share |")
    w("|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        w(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    w()

# ---- full captures
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "lts__t_sectors_op_red.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.sum",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
for kern in ("render_bwd", "render_fwd", "bin_scatter", "bin_count", "bin_prefix", "preprocess_bwd", "pack", "duplicate",
             "preprocess", "sort", "adam", "ssim_fwd", "ssim_bwd", "ref_render_bwd", "preprocess_bwd_gather"):
    rep = os.path.join(G, f"{tag}_{kern}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    w(f"## {kern}: ncu --set full (one launch)")
    w()
    w("| metric | value | unit |")
    w("|---|---:|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            w(f"| {k} | {vals[i]} | {units[i]} |")
    w()
out.close()
