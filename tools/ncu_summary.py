#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, stall reasons, SASS opcode mix.
usage: ncu_summary.py <report.ncu-rep> [kernel-regex] > profiles/<name>.md"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "lqr_"

def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units = raw[0], raw[1]
rows = [r for r in raw[2:] if re.search(kre, r[hdr.index("Kernel Name")])]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# ncu summary of `{rep}` (kernels matching `{kre}`)\n")
for r in rows:
    print(f"## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
    st = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
          if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h and r[i]]
    print("\nwarp stall reasons (warps per issue-active cycle): " +
          ", ".join(f"{h.split('issue_stalled_')[1].split('_per_')[0]} {v:.2f}" for v, h in sorted(st, reverse=True)[:8]) + "\n")
src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv"))))
h2 = None
byop, samp, first = collections.Counter(), collections.Counter(), True
for r in src:
    if r and r[0] == "Kernel Name":
        if h2 is not None and byop:
            break
        continue
    if r and r[0] == "Address":
        h2 = r
        continue
    if h2 is None or len(r) < len(h2):
        continue
    try:
        e, s = int(r[h2.index("Instructions Executed")] or 0), int(r[h2.index("# Samples")] or 0)
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h2.index("Source")])
    op = m.group(2) if m else "?"
    if not op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "SYNCS", "FFMA", "UBLKCP")):
        op = op.split(".")[0]
    byop[op] += e
    samp[op] += s
tot, ts = sum(byop.values()), max(1, sum(samp.values()))
print(f"## SASS opcode mix of the first matching kernel ({tot} warp instructions executed)\n")
print("| opcode | executed | share | stall samples |\n|---|---|---|---|")
for op, c in byop.most_common(24):
    print(f"| {op} | {c} | {100 * c / tot:.1f}% | {100 * samp[op] / ts:.1f}% |")
