import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
import bench
D = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()
U = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs()
eng = R.Engine(0); eng.set_timing(True)
n = 300000
ca, oa, cb, ob = bench.gen_pairs(n, 1, 4)
cells = float(((oa[1:] - oa[:-1]) * (ob[1:] - ob[:-1])).sum())
for label, costs, mut, force in [("ACGU user i16x2", U, None, 0), ("ACGU user forced i32", U, None, 2), ("ACGU user forced f64", U, None, 3),
                                 ("ACGU+N default (i32 x4)", D, 14, 0), ("15-letter default (f64)", D, "all", 0)]:
    a, b = ca.copy(), cb.copy()
    if mut == 14:
        rng = np.random.default_rng(2); a[rng.random(a.shape[0]) < 1e-3] = 14; b[rng.random(b.shape[0]) < 1e-3] = 14
    if mut == "all":
        rng = np.random.default_rng(2); a = rng.integers(0, 15, size=a.shape[0], dtype=np.uint8); b = rng.integers(0, 15, size=b.shape[0], dtype=np.uint8)
    eng.set_costs(costs)
    A, B = R.pack((a, oa)), R.pack((b, ob))
    for r in range(3):
        eng.distance_batch(A, B, force_mode=force)
    print(f"{label:28s} mode {eng.last_mode} kernel {eng.last_kernel_ms():8.3f} ms  {cells / eng.last_kernel_ms() * 1e-6:9.1f} GCUPS", flush=True)
