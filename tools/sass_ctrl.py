#!/usr/bin/env python
"""Decode the scheduling control bits of a SASS range (sm_100a, 128-bit encoding): stall count, yield, write / read
barrier, wait mask.  usage: tools/sass_ctrl.py <lib.so> <function substring> <addr_lo hex> <addr_hi hex>
The sum of the stall counts of a loop body is its issue time when no scoreboard wait bites."""
import re, subprocess, sys
so, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    lines = b.split("\n")
    total = 0; n = 0
    for i, l in enumerate(lines):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", l)
        if not m:
            continue
        addr = int(m.group(1), 16)
        if addr < lo or addr > hi:
            continue
        m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        hi64 = int(m2.group(1), 16)
        ctrl = hi64 >> 41
        stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wbar = (ctrl >> 5) & 7; rbar = (ctrl >> 8) & 7; wait = (ctrl >> 11) & 0x3f
        total += stall; n += 1
        print(f"{addr:05x} st={stall:2d} y={yld} wb={wbar if wbar != 7 else '-'} rb={rbar if rbar != 7 else '-'} wait={wait:06b}  {m.group(2).strip()}")
    print(f"== {name}: {n} instructions, sum of stalls {total}")
    break
