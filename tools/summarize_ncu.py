#!/usr/bin/env python3
"""Summarise gpurun_out/*.ncu-rep (ncu --set full) and launches.csv into profiles/<name>.md (tracked)."""
import csv
import subprocess
import sys
import collections

rep, launches, out, title = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]
lines = [f"# {title}", ""]
if rep != "-":
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines += [f"## ncu --set full --clock-control none ({rep.split('/')[-1]})", ""]
    for r in rows[2:]:
        lines.append(f"### {r[hdr.index('Kernel Name')][:100]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for w in WANT:
            if w in hdr:
                lines.append(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |")
        lines.append("")
        # warp states: warps per issue slot that sat in each stall reason (their sum = warp cycles per issued instruction)
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        st = sorted(((float(r[i].replace(",", "")), h[len(pre):-len(suf)]) for i, h in enumerate(hdr) if h.startswith(pre) and h.endswith(suf) and r[i]),
                    reverse=True)
        if st:
            lines.append("| warp state (per issued instruction) | warps |")
            lines.append("|---|---|")
            for v, k in st[:9]:
                lines.append(f"| {k} | {v:.2f} |")
            lines.append("")
        pipes = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active") and r[i]
                 and float(r[i].replace(",", "")) >= 1.0]
        if pipes:
            lines.append("| pipe (instructions, % of peak while active) | % |")
            lines.append("|---|---|")
            for h, v in sorted(pipes, key=lambda t: -float(t[1].replace(",", ""))):
                lines.append(f"| {h[len('sm__inst_executed_pipe_'):-len('.avg.pct_of_peak_sustained_active')]} | {float(v.replace(',', '')):.1f} |")
            lines.append("")
if launches != "-":
    agg = collections.OrderedDict()
    with open(launches) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        for r in rd:
            name = r[ki].split("(")[0][-70:]
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines += ["## launch list (ncu --metrics gpu__time_duration.sum, cold-cache serialised: compare shares)", "",
              "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f"| {k} | {n} | {t / 1e3:.1f} | {100 * t / tot:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
