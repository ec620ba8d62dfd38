#!/usr/bin/env python
"""Static SASS statistics per kernel of a built .so (cuobjdump -sass): instruction count and opcode histogram."""
import collections, re, subprocess, sys
so = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ""
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None; funcs = collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); funcs[fn] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and fn: funcs[fn].append(m.group(1))
for fn, ins in funcs.items():
    if pat and pat not in fn: continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
    print(fn, len(ins), "instrs; const-bank operand uses:", sum("c[0x3]" in i for i in ins))
    print("   ", ", ".join("%s %d" % kv for kv in ops.most_common(18)))
    if "--dump" in sys.argv:
        for i in ins: print("      ", i)
