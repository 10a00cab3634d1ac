import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import cost_tables
from _synth import c4_pair
eng = R.Engine(0); eng.set_costs(cost_tables.default_costs()); eng.set_timing(True)
for L, K in ((10000, 200), (20000, 100), (5000, 400)):
    pairs = [c4_pair(seed=9000 + k, m=L) for k in range(K)]
    cells = sum(float(a.shape[0]) * b.shape[0] for a, b in pairs)
    for rep in range(3):
        t0 = time.perf_counter(); out = eng.long_pairs(pairs); t = time.perf_counter() - t0; dev = eng.last_kernel_ms()
    print(f"long_pairs  {K} x {L}: wall {t*1e3:.1f} ms, device {dev:.1f} ms = {cells/dev*1e-6:.0f} GCUPS", flush=True)
    oa = np.zeros(K + 1, np.int64); ob = np.zeros(K + 1, np.int64)
    np.cumsum([a.shape[0] for a, _ in pairs], out=oa[1:]); np.cumsum([b.shape[0] for _, b in pairs], out=ob[1:])
    A = R.pack((np.concatenate([a for a, _ in pairs]), oa)); B = R.pack((np.concatenate([b for _, b in pairs]), ob))
    for rep in range(3):
        t0 = time.perf_counter(); res = eng.script_batch(A, B); t = time.perf_counter() - t0; dev = eng.last_kernel_ms()
    ok = all(np.array_equal(res["op"][p, :res["n_ops"][p]], out[p]["op"]) for p in range(0, K, 17))
    for rep in range(3):
        t0 = time.perf_counter(); d = eng.distance_batch(A, B); td = time.perf_counter() - t0; devd = eng.last_kernel_ms()
    for rep in range(2):
        t0 = time.perf_counter(); outd = eng.long_pairs(pairs, want_script=False); tl = time.perf_counter() - t0; devl = eng.last_kernel_ms()
    okd = all(d[p] == outd[p]["dist"] for p in range(K))
    print(f"distance_batch {K} x {L}: wall {td*1e3:.1f} ms, kernel {devd:.1f} ms = {cells/devd*1e-6:.0f} GCUPS (mode {eng.last_mode}); long_pairs distance only: wall {tl*1e3:.1f} ms, device {devl:.1f} ms = {cells/devl*1e-6:.0f} GCUPS; equal: {okd}", flush=True)
    print(f"script_batch {K} x {L}: wall {t*1e3:.1f} ms, device {dev:.1f} ms = {cells/dev*1e-6:.0f} GCUPS; same scripts: {ok}", flush=True)
