#!/usr/bin/env python
"""Copy the files of the unmodified reference that its Wagner-Fischer path needs into baseline/_ref/ (git-ignored,
NOT gpurun-ignored: it travels to the GPU box with the snapshot).  bench.py times that copy
— `StringEditDistance.wagnerFisher` itself, pure Python — next to the GPU path (`cpu_baseline_python`).
Run in the build container, where /root/reference exists; __graft_entry__.build() calls it."""
import os
import shutil
import sys

REF = os.environ.get("RSD_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["StringEditDistance.py", "costs.json", "user_costs.json"]


def make(verbose=False) -> bool:
    if not all(os.path.exists(os.path.join(REF, f)) for f in FILES):
        return os.path.exists(os.path.join(DST, FILES[0]))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        if not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst):
            shutil.copyfile(src, dst)
            if verbose:
                print("copied", src, "->", dst, file=sys.stderr)
    return True


if __name__ == "__main__":
    print(make(verbose=True))
