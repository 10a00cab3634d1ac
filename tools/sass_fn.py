#!/usr/bin/env python
"""Print the SASS of the functions whose mangled name contains PATTERN, with an opcode histogram.
usage: tools/sass_fn.py <lib.so> <pattern> [--hist-only] [--loop]"""
import re, subprocess, sys, collections
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if pat not in name:
        continue
    lines = [l for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4}\*/", l)]
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", l)
        if m: ops[m.group(1)] += 1
    print("==", name, len(lines), "instructions")
    if "--hist-only" not in sys.argv:
        print("\n".join(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip() for l in lines))
    for k, v in ops.most_common(40):
        print(f"  {v:5d} {k}")
