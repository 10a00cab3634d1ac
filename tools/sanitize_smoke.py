#!/usr/bin/env python
"""Small run of every C-ABI entry point, meant to be wrapped in compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from oracle import oracle as O

D = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()
U = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).user_costs()
rng = np.random.default_rng(3)
def seqs(n, lo, hi, al):
    al = np.array(list(al)); return ["".join(al[rng.integers(0, len(al), size=L)]) for L in rng.integers(lo, hi + 1, size=n)]
eng = R.Engine(0)
for costs, al in ((U, "AGCU"), (D, "AGCUN"), (D, R.SYMBOLS)):
    eng.set_costs(costs)
    a = seqs(300, 0, 330, al) + seqs(6, 1000, 1100, al); b = seqs(300, 0, 330, al) + seqs(6, 1000, 2100, al)
    got = eng.distance_batch(R.pack(a), R.pack(b))
    ac, ao = O.concat(a); bc, bo = O.concat(b)
    assert np.array_equal(got, O.distance_batch(ac, ao, bc, bo, costs)), "distance"
    res = eng.script_batch(R.pack(a[:150]), R.pack(b[:150]), check_roundtrip=True)
    assert res["ok"].all()
    out, ol, err = eng.patch_batch(res, R.pack(a[:150]), R.pack(b[:150]), R.pack(a[:150]))
    assert (err == 0).all()
    print("ok", al, "mode", eng.last_mode, flush=True)
eng.set_costs(D)
v, m = eng.matrix(O.encode("AGRGAUUCG"), O.encode("AGGGAACG"))
recs = seqs(5000, 24, 31, "AGCUN"); eng.db_load(R.pack(recs, bits=4))
i1, s1, alls = eng.db_search_topk(R.pack(seqs(5, 24, 31, "AGCU"), bits=4), 10, want_scores=True)
i2, s2 = eng.db_search_topk(R.pack(seqs(5, 24, 31, "AGCU"), bits=4), 10, force_mode=3)
eng.db_free()
a = rng.integers(0, 4, size=3000, dtype=np.uint8); b = rng.integers(0, 4, size=2777, dtype=np.uint8)
r = eng.long_pair(a, b); r2 = eng.long_pair(a, b, force_mode=3)
assert r["dist"] == r2["dist"] and np.array_equal(r["op"], r2["op"])
print("sanitize smoke done", eng.launch_count(), "launches")
