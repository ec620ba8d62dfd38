#!/usr/bin/env python
"""Aggregate an ncu report's per-SASS-instruction samples / executed counts by CUDA source line.
ncu's CSV source page is SASS-only, so line info is taken from `nvdisasm --print-line-info` of the cubin
extracted from the .so and joined on the instruction offset within the kernel."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, so = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
kname = rows[0][1]
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ci, si, sm, ai = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Address")
ins = [(int(r[ai], 16), r[si].strip(), int(r[ci]) if r[ci].isdigit() else 0, int(r[sm]) if r[sm].isdigit() else 0)
       for r in rows[hi + 1:] if len(r) > ci and r[ai].startswith("0x")]
base = ins[0][0]
# mangled name from demangled: find via cuobjdump listing
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
best = None
for cb in cubins:
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", cb], capture_output=True, text=True).stdout
    # split per function
    cur = None; line = None; funcs = {}
    for l in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = []; line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2)), "inlined" in m.group(3)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur is not None: funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), line))
    for fn, lst in funcs.items():
        if len(lst) == len(ins):
            dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()
            if dem.replace(" ", "")[:40] == kname.replace("(int)", "").replace("(bool)", "").replace(" ", "")[:40] or best is None:
                best = lst
if best is None:
    sys.exit("could not match kernel to a cubin function (%d instrs)" % len(ins))
agg = collections.defaultdict(lambda: [0, 0])
tot_s = sum(i[3] for i in ins); tot_n = sum(i[2] for i in ins)
for (addr, sass, n, smp), (off, sass2, line) in zip(ins, best):
    key = "%s:%d" % (line[0], line[1]) if line else "?"
    agg[key][0] += smp; agg[key][1] += n
print(kname)
print("total samples %d, warp instructions %d" % (tot_s, tot_n))
print("%-34s %8s %6s %12s %6s" % ("line", "samples", "%", "warp-instr", "%"))
for key, (smp, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-34s %8d %5.1f%% %12d %5.1f%%" % (key, smp, 100 * smp / max(tot_s, 1), n, 100 * n / max(tot_n, 1)))
