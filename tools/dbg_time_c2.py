"""Kernel time of the C2 distance batch without any result check (for instruction-mix experiments)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import rna_sequence_diff_patch_b200 as R
import bench
ca, oa, cb, ob = bench.gen_pairs(1_000_000, bench.SEEDS["c2"], 4)
A = R.pack((ca, oa)); B = R.pack((cb, ob))
eng = R.Engine(0); eng.set_costs(__import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs()); eng.set_timing(True)
ms = []
for r in range(6):
    eng.distance_batch(A, B)
    ms.append(eng.last_kernel_ms())
print("kernel ms (sum of chunk kernels):", [round(x, 3) for x in ms])
