import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import cost_tables
from _synth import c4_pair
eng = R.Engine(0); eng.set_costs(cost_tables.default_costs()); eng.set_timing(True)
for L, K in ((1500, 4000), (2000, 3000), (3000, 1500), (4000, 1000)):
    pairs = [c4_pair(seed=7000 + k, m=L) for k in range(K)]
    cells = sum(float(a.shape[0]) * b.shape[0] for a, b in pairs)
    oa = np.zeros(K + 1, np.int64); ob = np.zeros(K + 1, np.int64)
    np.cumsum([a.shape[0] for a, _ in pairs], out=oa[1:]); np.cumsum([b.shape[0] for _, b in pairs], out=ob[1:])
    A = R.pack((np.concatenate([a for a, _ in pairs]), oa)); B = R.pack((np.concatenate([b for _, b in pairs]), ob))
    for name, env in (("tape ", {"RSD_SCRIPT_NO_LONG": "1", "RSD_DIST_NO_LONG": "1"}), ("panel", {"RSD_SCRIPT_LONG_MIN": "1", "RSD_DIST_LONG_MIN": "1"})):
        os.environ.update(env)
        for rep in range(3):
            t0 = time.perf_counter(); res = eng.script_batch(A, B); ts = time.perf_counter() - t0; ds = eng.last_kernel_ms()
        for rep in range(3):
            t0 = time.perf_counter(); d = eng.distance_batch(A, B); td = time.perf_counter() - t0; dd = eng.last_kernel_ms()
        for k in env: del os.environ[k]
        print(f"{name} {K} x {L}: script wall {ts*1e3:.1f} ms device {ds:.1f} ms = {cells/ds*1e-6:.0f} GCUPS | distance wall {td*1e3:.1f} ms kernel {dd:.1f} ms = {cells/dd*1e-6:.0f} GCUPS", flush=True)
