#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics, opcode mix and hot SASS blocks.
usage: ncu_summary.py report.ncu-rep [out.txt]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout


def P(*a):
    print(*a, file=out)


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__lsu_writeback_active_mem_lg.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
        "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_membar",
        "smsp__pcsamp_warps_issue_stalled_sleeping",
        "smsp__pcsamp_warps_issue_stalled_drain",
        "smsp__pcsamp_warps_issue_stalled_imc_miss",
        "smsp__pcsamp_sample_buffer_full"]
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in want or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
            if h.startswith("smsp__pcsamp") and r[i] in ("0", ""):
                continue
            P("%-70s %-10s %s" % (h, rows[1][i], r[i]))
    P("-" * 60)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# may hold several kernels; take the first table
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iN]) for r in data)
P("total warp instructions", tot, "sass lines", len(data), "samples", tots)
ops, samp = collections.Counter(), collections.Counter()


def opname(s):
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    base = op.split(".")[0]
    if ".MOV" in op:
        base += ".MOV"
    return base


for r in data:
    ops[opname(r[iS])] += int(r[iE])
    samp[opname(r[iS])] += int(r[iN])
for op, c in ops.most_common(22):
    P("%-12s %12d %5.1f%%  samples %5.1f%%" % (op, c, 100.0 * c / tot,
                                              100.0 * samp[op] / max(1, tots)))
blocks, cur = [], None
for idx, r in enumerate(data):
    e = int(r[iE])
    if cur and cur[0] == e:
        cur[2] = idx
    else:
        cur = [e, idx, idx]
        blocks.append(cur)
P("hot blocks (contiguous SASS with equal execution count):")
for e, a, b in sorted(blocks, key=lambda x: -x[0] * (x[2] - x[1] + 1))[:10]:
    n = b - a + 1
    c = collections.Counter(opname(r[iS]) for r in data[a:b + 1])
    s = sum(int(r[iN]) for r in data[a:b + 1])
    P("  sass[%d:%d] n=%d exec=%d inst-share=%.1f%% sample-share=%.1f%% %s"
      % (a, b, n, e, 100.0 * e * n / tot, 100.0 * s / max(1, tots),
         dict(c.most_common(7))))
