#!/usr/bin/env python
"""Split an ncu report's SASS listing at BAR.SYNC instructions and report samples / executed warp-instructions per
segment (kernel phases are separated by __syncthreads, and code is laid out in program order)."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ci, si, sm = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
seg = 0; segs = collections.OrderedDict(); first = {}
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= ci or not r[ci].isdigit(): continue
    s = r[si].strip()
    d = segs.setdefault(seg, {"smp": 0, "n": 0, "ops": collections.Counter(), "stalls": collections.Counter(), "start": idx})
    n = int(r[ci]); sv = int(r[sm]) if r[sm].isdigit() else 0
    d["smp"] += sv; d["n"] += n
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
    d["ops"][op] += n
    for c in stall_cols:
        if r[c].isdigit(): d["stalls"][hdr[c]] += int(r[c])
    if "BAR.SYNC" in s or s.startswith("EXIT") or " EXIT" in s and "@" not in s: seg += 1
tot_s = sum(d["smp"] for d in segs.values()); tot_n = sum(d["n"] for d in segs.values())
for k, d in segs.items():
    if d["n"] == 0 and d["smp"] == 0: continue
    print("segment %2d (sass #%5d..): samples %7d (%4.1f%%)  warp-instr %11d (%4.1f%%)  top ops: %s" % (
        k, d["start"], d["smp"], 100 * d["smp"] / tot_s, d["n"], 100 * d["n"] / tot_n,
        ", ".join("%s %.0f%%" % (o, 100 * c / max(d["n"], 1)) for o, c in d["ops"].most_common(6))))
    if d["stalls"]:
        ts = sum(d["stalls"].values())
        print("            stalls: " + ", ".join("%s %.0f%%" % (o.replace("stall_", ""), 100 * c / ts) for o, c in d["stalls"].most_common(6)))
