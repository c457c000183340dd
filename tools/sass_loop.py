#!/usr/bin/env python
"""Lists the backward-branch loops of one kernel in a cuobjdump -sass dump:
   cuobjdump -sass x.o | tools/sass_loop.py <substring of the mangled name>
For each loop: start, end, instruction slots, and a histogram of opcodes."""
import collections
import re
import sys

pat = sys.argv[1]
ins = []
on = False
for line in sys.stdin:
    if "Function :" in line:
        if on:
            break
        on = pat in line
        continue
    if on:
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
for a, t in ins:
    m = re.search(r"BRA\S* (?:.*, )?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        body = [x for x in ins if lo <= x[0] <= a]
        h = collections.Counter(re.sub(r"^@!?U?P\w+ ", "", x[1]).split()[0].split(".")[0] for x in body)
        print(f"loop {lo:#x}..{a:#x}: {len(body)} slots  " + " ".join(f"{k}:{v}" for k, v in h.most_common(14)))
