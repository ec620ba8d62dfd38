#!/usr/bin/env python
"""Print the SASS of one kernel between two addresses. usage: sass_dump.py <lib.so> <kernel-substring> <lo-hex> <hi-hex>"""
import re, subprocess, sys
so, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None
for l in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m: cur = m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m and cur and pat in cur and lo <= int(m.group(1), 16) <= hi:
        print(m.group(1), m.group(2))
