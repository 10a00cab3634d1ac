import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()); eng.set_timing(True)
rng = np.random.default_rng(1)
for m, n in [(50000, 256), (50000, 512), (50000, 1024), (50000, 4096), (50000, 16384), (50000, 50000), (10000, 50000), (2000, 50000)]:
    a = rng.integers(0, 4, size=m, dtype=np.uint8); b = rng.integers(0, 4, size=n, dtype=np.uint8)
    for r in range(2):
        eng.long_pair(a, b, want_script=False)
    ms = eng.last_kernel_ms()
    print(f"m={m} n={n} panels={(n+255)//256} fwd_ms={ms:.3f} ns/step={ms*1e6/(m+31):.1f} gcups={m*n/ms*1e-6:.1f}", flush=True)
