"""GPU parity: query batch vs database with exact top-k (score desc, index asc)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as G
    G.build()
    import rna_sequence_diff_patch_b200 as R
    assert R.load_library().rsd_device_count() > 0
    return R


@pytest.fixture(scope="module")
def eng(R):
    return R.Engine(0)


def make_db(rng, n, lo=24, hi=31, p_n=1e-3, iupac_frac=0.0):
    lens = rng.integers(lo, hi + 1, size=n)
    off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    codes[rng.random(codes.shape[0]) < p_n] = 14                      # N
    if iupac_frac > 0:
        for r in np.nonzero(rng.random(n) < iupac_frac)[0]:
            codes[off[r]:off[r + 1]] = rng.integers(0, 15, size=lens[r], dtype=np.uint8)
    return codes, off


def make_queries(rng, codes, off, nq, rate=0.1):
    qs = []
    for r in rng.integers(0, len(off) - 1, size=nq):
        s = codes[off[r]:off[r + 1]].copy()
        hit = rng.random(s.shape[0]) < rate
        s[hit] = rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)
        qs.append(O.decode(s))
    return qs


def oracle_topk(queries, codes, off, costs, k):
    out_i, out_s, alls = [], [], []
    for q in queries:
        i, s, a = O.search_topk(q, codes, off, costs, k, want_scores=True)
        out_i.append(i); out_s.append(s); alls.append(a)
    return np.stack(out_i), np.stack(out_s), np.stack(alls)


def test_fast_path_acgun_topk_and_all_scores(R, eng, golden):
    rng = np.random.default_rng(20260005)
    codes, off = make_db(rng, 40000)
    queries = make_queries(rng, codes, off, 6)
    costs = golden["default_costs"]
    eng.set_costs(costs)
    db = R.pack((codes, off), bits=4)
    eng.db_load(db)
    idx, sc, alls = eng.db_search_topk(R.pack(queries, bits=4), 10, want_scores=True)
    assert eng.last_mode == 1, "ACGU+N under default costs must take the int16x2 search kernel"
    wi, ws, wa = oracle_topk(queries, codes, off, costs, 10)
    assert np.array_equal(alls, wa)
    assert np.array_equal(idx, wi) and np.array_equal(sc, ws)
    idx2, sc2 = eng.db_search_topk(R.pack(queries, bits=4), 10, force_mode=3)      # general fp64 path agrees
    assert np.array_equal(idx2, wi) and np.array_equal(sc2, ws)
    eng.db_free()


def test_ties_keep_collection_order_and_short_db(R, eng, golden):
    base = ["ACGUACGUACGUACGUACGUACGUAC", "ACGUACGUACGUACGUACGUACGUAG", "GGGGGGGGGGGGGGGGGGGGGGGGGG"]
    seqs = [base[k % 3] for k in range(3000)]
    eng.set_costs(golden["default_costs"])
    c, o = O.concat(seqs)
    eng.db_load(R.pack(seqs, bits=4))
    q = ["ACGUACGUACGUACGUACGUACGUAC", "GGGGGGGGGGGGGGGGGGGGGGGGGA"]
    idx, sc = eng.db_search_topk(R.pack(q, bits=4), 25)
    wi, ws, _ = oracle_topk(q, c, o, golden["default_costs"], 25)
    assert np.array_equal(idx, wi) and np.array_equal(sc, ws)
    assert idx[0].tolist() == list(range(0, 75, 3))                   # equal scores come out in collection order
    eng.db_free()
    eng.db_load(R.pack(seqs[:4], bits=4))
    idx, sc = eng.db_search_topk(R.pack(q, bits=4), 10)
    assert idx[0, :4].tolist() == [0, 3, 1, 2] and (idx[:, 4:] == -1).all()
    eng.db_free()


def test_iupac_records_general_fp64_path(R, eng, golden):
    rng = np.random.default_rng(32)
    codes, off = make_db(rng, 6000, iupac_frac=0.01)
    queries = make_queries(rng, codes, off, 3)
    for costs in (golden["default_costs"], golden["user_costs"]):
        eng.set_costs(costs)
        eng.db_load(R.pack((codes, off), bits=4))
        idx, sc, alls = eng.db_search_topk(R.pack(queries, bits=4), 10, want_scores=True)
        assert eng.last_mode == 3
        wi, ws, wa = oracle_topk(queries, codes, off, costs, 10)
        assert np.array_equal(alls, wa) and np.array_equal(idx, wi) and np.array_equal(sc, ws)
        eng.db_free()


def test_long_records_general_path(R, eng, golden):
    rng = np.random.default_rng(33)
    codes, off = make_db(rng, 3000, lo=20, hi=90, p_n=0.0)
    queries = make_queries(rng, codes, off, 3)
    eng.set_costs(golden["user_costs"])
    eng.db_load(R.pack((codes, off), bits=2))
    idx, sc = eng.db_search_topk(R.pack(queries, bits=2), 7)
    wi, ws, _ = oracle_topk(queries, codes, off, golden["user_costs"], 7)
    assert np.array_equal(idx, wi) and np.array_equal(sc, ws)
    eng.db_free()


def test_sharded_search_equals_single_shard(R, eng, golden):
    """Ranks emulated one after another on one GPU (no co-resident ranks): shard, local top-k,
    merge == the unsharded answer; also crosses the 32768 / 1M chunk boundaries."""
    from rna_sequence_diff_patch_b200.dist_search import shard_bounds, slice_packed
    from rna_sequence_diff_patch_b200.engine import topk_merge
    rng = np.random.default_rng(34)
    codes, off = make_db(rng, 1_100_000)
    queries = make_queries(rng, codes, off, 2)
    eng.set_costs(golden["default_costs"])
    db = R.pack((codes, off), bits=4)
    Q = R.pack(queries, bits=4)
    eng.db_load(db)
    gi, gs = eng.db_search_topk(Q, 10)
    eng.db_free()
    wi, ws, _ = oracle_topk(queries, codes, off, golden["default_costs"], 10)
    assert np.array_equal(gi, wi) and np.array_equal(gs, ws)
    for world in (2, 3, 8):
        li, ls = [], []
        for lo, hi in shard_bounds(db.len, world):
            eng.db_load(slice_packed(db, lo, hi), global_index_base=lo)
            i, s = eng.db_search_topk(Q, 10)
            li.append(i); ls.append(s)
            eng.db_free()
        mi, ms = topk_merge(np.stack(li), np.stack(ls))
        assert np.array_equal(mi, gi) and np.array_equal(ms, gs), world


@pytest.mark.parametrize("nq,n_db,k", [(64, 120_000, 10), (130, 120_000, 10), (66, 1_100_000, 7), (8, 2_200_000, 7)])
def test_c5_shaped_query_batches_every_query_vs_oracle(R, eng, golden, nq, n_db, k):
    """BASELINE config 5 at its stated shape: a full batch of 64 queries (one QB batch, seed chunk spread over
    gridDim.y), 130 queries (three batches, the last one partial), a database of more than 2^20 records and one that crosses
    the 2^21-record chunk boundary — every query's top-k against the oracle (IR:466-477 + performance.py:12-15)."""
    rng = np.random.default_rng(20260005 + nq)
    codes, off = make_db(rng, n_db)
    queries = make_queries(rng, codes, off, nq)
    costs = golden["default_costs"]
    eng.set_costs(costs)
    eng.db_load(R.pack((codes, off), bits=4))
    try:
        idx, sc = eng.db_search_topk(R.pack(queries, bits=4), k)
        assert eng.last_mode == 1
    finally:
        eng.db_free()
    for q in range(nq):
        wi, ws = O.search_topk(queries[q], codes, off, costs, k)
        assert np.array_equal(idx[q], wi) and np.array_equal(sc[q], ws), q


def test_multi_device_search_one_process(R, golden):
    """rsd_multi_*: shards on several contexts of ONE process, gather, merge == the oracle; all_scores in record
    order.  With one GPU in the box the device is listed three times (emulated shards, gather by device copies);
    with several GPUs the real thing runs: one context per device and a grouped ncclAllGather."""
    rng = np.random.default_rng(35)
    codes, off = make_db(rng, 90_000)
    queries = make_queries(rng, codes, off, 5)
    costs = golden["default_costs"]
    n_dev = R.load_library().rsd_device_count()
    for devices in ([0, 0, 0], None if n_dev > 1 else [0]):
        me = R.MultiEngine(devices)
        try:
            me.set_costs(costs)
            me.db_load(R.pack((codes, off), bits=4))
            idx, sc, alls = me.db_search_topk(R.pack(queries, bits=4), 10, want_scores=True)
            assert me.last_mode == 1
            wi, ws, wa = oracle_topk(queries, codes, off, costs, 10)
            assert np.array_equal(idx, wi) and np.array_equal(sc, ws) and np.array_equal(alls, wa)
            idx2, sc2 = me.db_search_topk(R.pack(queries[:2], bits=4), 3)
            assert np.array_equal(idx2, wi[:2, :3]) and np.array_equal(sc2, ws[:2, :3])
            me.db_free()
            # fewer records than devices: empty shards pad with -1 and the merge skips them
            me.db_load(R.pack((codes[:off[2]], off[:3].copy()), bits=4))
            idx3, sc3 = me.db_search_topk(R.pack(queries[:1], bits=4), 4)
            w3 = O.search_topk(queries[0], codes[:off[2]], off[:3].copy(), costs, 4)
            assert np.array_equal(idx3[0], w3[0]) and np.array_equal(sc3[0][:2], w3[1][:2])
        finally:
            me.close()


def test_mixed_database_fast_prefix_plus_general_rest(R, eng, golden):
    """SURVEY 8d, second C5 run: 1 % of the records drawn from all 15 symbols.  The store order puts the A/G/C/U(+N)
    records first (tiers by rarest symbol, rsd_db_load); the int16x2 kernel scores that prefix, the fp64 kernels the
    IUPAC records (0.66 / 0.83 are not dyadic) — every score, and the top-k with its collection-order tie-break, must
    equal the oracle's, also for a query that itself carries an ambiguity code."""
    rng = np.random.default_rng(515)
    codes, off = make_db(rng, 30000, iupac_frac=0.01)
    queries = make_queries(rng, codes, off, 5)
    q_amb = list(queries[0]); q_amb[3] = "R"; queries.append("".join(q_amb))   # puts every record on the general path for that call
    costs = golden["default_costs"]
    eng.set_costs(costs)
    eng.db_load(R.pack((codes, off), bits=4))
    try:
        launches0 = eng.launch_count()
        idx, sc, alls = eng.db_search_topk(R.pack(queries[:5], bits=4), 10, want_scores=True)
        n_launch = eng.launch_count() - launches0
        assert eng.last_mode == 3                       # the mode of the slowest part
        wi, ws, wa = oracle_topk(queries[:5], codes, off, costs, 10)
        assert np.array_equal(alls, wa)
        assert np.array_equal(idx, wi) and np.array_equal(sc, ws)
        # the fast kernel took the prefix: far fewer launches than 3 per query per chunk over the whole database
        assert n_launch < 60
        idx, sc, alls = eng.db_search_topk(R.pack(queries[5:], bits=4), 10, want_scores=True)
        wi, ws, wa = oracle_topk(queries[5:], codes, off, costs, 10)
        assert np.array_equal(alls, wa) and np.array_equal(idx, wi) and np.array_equal(sc, ws)
        # the ambiguous query in the middle of a batch: the library groups the queries (fast-prefix ones first) and
        # returns the rows in the caller's order
        mixed = [queries[1], queries[5], queries[0], queries[5], queries[3]]
        launches0 = eng.launch_count()
        idx, sc, alls = eng.db_search_topk(R.pack(mixed, bits=4), 10, want_scores=True)
        assert eng.launch_count() - launches0 < 700    # 2 ambiguous queries x ~100 launches, not 5 x
        wi, ws, wa = oracle_topk(mixed, codes, off, costs, 10)
        assert np.array_equal(alls, wa) and np.array_equal(idx, wi) and np.array_equal(sc, ws)
        # user costs: every symbol pair of this table is dyadic? (A->K 1.5 ...) -> whatever the split, results equal the oracle
        eng.set_costs(golden["user_costs"])
        idx, sc, alls = eng.db_search_topk(R.pack(queries[:3], bits=4), 7, want_scores=True)
        wi, ws, wa = oracle_topk(queries[:3], codes, off, golden["user_costs"], 7)
        assert np.array_equal(alls, wa) and np.array_equal(idx, wi) and np.array_equal(sc, ws)
    finally:
        eng.db_free()
