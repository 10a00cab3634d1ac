"""C5 through ONE process driving every visible GPU (rsd_multi_*, MultiEngine: what dropin/IRMethods.search_collection
uses): contiguous shards, the query batch on every device, local top-k, one ncclAllGather, merge.  Wall time per
64-query batch over 10^7 records, checked against the oracle for two queries."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as G; G.build()
import rna_sequence_diff_patch_b200 as R
from rna_sequence_diff_patch_b200 import cost_tables
from oracle import oracle as O

n_rec = int(os.environ.get("RECORDS", 10_000_000)); nq, k = 64, 10
rng = np.random.default_rng(20260005)
lens = rng.integers(24, 32, size=n_rec)
off = np.zeros(n_rec + 1, np.int64); np.cumsum(lens, out=off[1:])
codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
codes[rng.random(codes.shape[0]) < 1e-3] = 14
qs, qo = [], [0]
for r in rng.integers(0, n_rec, size=nq):
    s = codes[off[r]:off[r + 1]].copy(); hit = rng.random(s.shape[0]) < 0.1
    s[hit] = rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)
    qs.append(s); qo.append(qo[-1] + s.shape[0])
Q = R.pack((np.concatenate(qs), np.array(qo, np.int64)), bits=4)
me = R.MultiEngine()
me.set_costs(cost_tables.default_costs())
t0 = time.perf_counter(); me.db_load(R.pack((codes, off), bits=4)); t_load = time.perf_counter() - t0
ts = []
for rep in range(6):
    t0 = time.perf_counter(); idx, sc = me.db_search_topk(Q, k); ts.append(time.perf_counter() - t0)
for qi in (0, nq - 1):
    wi, ws = O.search_topk(O.decode(qs[qi]), codes, off, cost_tables.default_costs(), k)
    assert np.array_equal(idx[qi], wi) and np.array_equal(sc[qi], ws), f"query {qi} differs from the oracle"
cells = float(np.array([len(q) for q in qs], np.float64).sum() * lens.astype(np.float64).sum())
t = float(np.mean(sorted(ts[1:])[:3]))
print(json.dumps({"config": "C5 one process, all GPUs (rsd_multi)", "devices": me.n_devices, "records": n_rec, "queries": nq, "k": k,
                  "wall_ms_per_batch": t * 1e3, "gcups_wall": cells / t * 1e-9, "db_load_s": t_load, "oracle_checked_queries": 2}))
