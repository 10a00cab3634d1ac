import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import rna_sequence_diff_patch_b200 as R
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()); eng.set_timing(True)
rng = np.random.default_rng(1)
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
a = rng.integers(0, 4, size=m, dtype=np.uint8); b = rng.integers(0, 4, size=n, dtype=np.uint8)
for r in range(2):
    eng.long_pair(a, b, want_script=False)
print("fwd ms", eng.last_kernel_ms())
