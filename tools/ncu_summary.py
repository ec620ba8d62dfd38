#!/usr/bin/env python
"""Summarise one .ncu-rep: headline metrics, opcode histogram, top stall sites (reads `ncu -i` CSV pages)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; branches = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warp") and "per_issue_active" in h and "_not_issued" not in h) or \
       (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print("%-95s %-12s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Thread Instructions Executed")
sm = hdr.index("Warp Stall Sampling (All Samples)")
ops = collections.Counter(); tot = 0; samples = []; alls = 0
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci].isdigit(): continue
    s = re.sub(r"^@!?U?P\d+\s+", "", r[si].strip()); op = s.split()[0].split(".")[0]
    n = int(r[ci]); ops[op] += n; tot += n
    sv = int(r[sm]) if r[sm].isdigit() else 0; alls += sv; samples.append((sv, n, r[0], r[si].strip()))
print("warp instructions executed:", tot, ("= %.1f thread-instr per branch" % (tot * 32 / branches)) if branches else "")
print("opcode mix:", ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in ops.most_common(24)))
samples.sort(reverse=True)
print("top stall-sample sites (samples, %% of all, executed, sass):")
for sv, n, addr, s in samples[:40]:
    print("  %7d %5.1f%% %10d  %s" % (sv, 100 * sv / max(alls, 1), n, s[:110]))
