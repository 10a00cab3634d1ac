"""Several long pairs in flight on one GPU: one context (own streams and buffers) and one host thread per pair.
The panel pipeline of a single 50 kb pair leaves most issue slots idle (one warp per scheduler, half the time in
pipeline fill); independent pairs fill them."""
import os, sys, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import cost_tables
from _synth import c4_pair
import torch
pairs = [c4_pair(seed=20260004 + k) for k in range(8)]
for K in (1, 2, 3, 4, 6, 8):
    engs = [R.Engine(0) for _ in range(K)]
    for e in engs: e.set_costs(cost_tables.default_costs())
    res = [None] * K
    def work(k, want):
        res[k] = engs[k].long_pair(*pairs[k], want_script=want)
    for want in (False, True):
        for rep in range(3):
            th = [threading.Thread(target=work, args=(k, want)) for k in range(K)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for t in th: t.start()
            for t in th: t.join()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        cells = sum(float(a.shape[0]) * b.shape[0] for a, b in pairs[:K])
        print(f"{K} pairs in flight, script={want}: {dt * 1e3:.2f} ms wall, {cells / dt * 1e-9:.0f} GCUPS aggregate", flush=True)
    for e in engs: e.close()
