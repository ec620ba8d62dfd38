#!/usr/bin/env python
"""Top-N SASS instructions of an ncu report by stall samples, with the dominant stall reasons and a few preceding
instructions for context.  usage: ncu_hot.py <report> [N] [context]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ci, si, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[hi + 1:]:
    if len(r) <= ci or not r[ci].isdigit(): continue
    st = {hdr[c].replace("stall_", ""): int(r[c]) for c in stall_cols if r[c].isdigit() and int(r[c]) > 0}
    ins.append((r[si].strip(), int(r[ci]), int(r[sm]) if r[sm].isdigit() else 0, st))
tot = sum(i[2] for i in ins)
agg = collections.Counter()
for s, n, smp, st in ins:
    for k, v in st.items(): agg[k] += v
print("total samples", tot, "| stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(sum(agg.values()), 1)) for k, v in agg.most_common(10)))
order = sorted(range(len(ins)), key=lambda i: -ins[i][2])[:N]
for i in order:
    s, n, smp, st = ins[i]
    for j in range(max(0, i - ctx), i):
        print("        . #%-5d %-60s x%d" % (j, ins[j][0][:60], ins[j][1]))
    print("%5.1f%% #%-5d %-60s x%-8d %s" % (100 * smp / tot, i, s[:60], n, ", ".join("%s %d" % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:3])))
