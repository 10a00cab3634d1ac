"""Long-pair (C4) timing and pipeline trace: RSD_TRACE prints per-panel start / end times; the slope of the end
times over the panel index is the achieved lag per panel."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import cost_tables
from _synth import c4_pair
a, b = c4_pair(m=int(os.environ.get("L", 50000)))
eng = R.Engine(0); eng.set_costs(cost_tables.default_costs()); eng.set_timing(True)
for r in range(4):
    eng.long_pair(a, b, want_script=False)
    print("forward ms", round(eng.last_kernel_ms(), 3), flush=True)
os.environ["RSD_TRACE"] = "1"
eng.long_pair(a, b, want_script=False)
