#!/usr/bin/env python
"""Static loop census of one kernel's SASS: for every backward branch prints the loop body size and opcode mix.
usage: sass_loops.py <lib.so> <kernel-name-substring>"""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None; funcs = collections.OrderedDict()
for l in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for fn, ins in funcs.items():
    if pat not in fn: continue
    print("==", fn, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, s) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", s)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                body = ins[addr[tgt]:i + 1]
                ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for _, x in body)
                f64 = sum(c for o, c in ops.items() if o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
                print("  loop %04x..%04x: %4d instr, %3d fp64 | %s" % (tgt, a, len(body), f64, ", ".join("%s %d" % oc for oc in ops.most_common(14))))
