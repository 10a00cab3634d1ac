"""Set / multiset / TF-vector similarity search on the GPU: rsd_db_similarity against the
reference's recorded outputs (tests/golden/ref_golden_ir.json) and the numpy oracle — bit for bit."""
import json
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ir_oracle as IO  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as G
    G.build()
    import rna_sequence_diff_patch_b200 as R
    return R


@pytest.fixture(scope="module")
def eng(R):
    return R.Engine(0)


@pytest.fixture(scope="module")
def gir():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_golden_ir.json")))


def unhex(xs):
    return np.array([math.nan if x == "nan" else float.fromhex(x) for x in xs], dtype=np.float64)


def same(a, b):
    return bool(np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64)))


@pytest.mark.parametrize("bits", [4])
def test_all_methods_match_reference_golden(R, eng, gir, bits):
    eng.db_load(R.pack(gir["docs"], bits=bits))
    try:
        for q in gir["queries"]:
            qc = R.encode(q)
            for method in IO.METHODS:
                got, _, _ = eng.db_similarity(qc, method)
                assert same(got, unhex(gir["scores"][q][method])), (q, method)
    finally:
        eng.db_free()


def test_two_bit_database_and_random_batches_match_oracle(R, eng):
    rng = np.random.default_rng(31)
    acgu = ["".join(rng.choice(list("AGCU"), size=int(L))) for L in rng.integers(1, 70, size=400)]
    iupac = ["".join(rng.choice(list(IO.SYMBOLS), size=int(L))) for L in rng.integers(1, 50, size=400)]
    for docs, bits in ((acgu, 2), (iupac, 4), (acgu + iupac, 4)):
        eng.db_load(R.pack(docs, bits=bits))
        try:
            for q in (docs[3], docs[17], "ACGUNNRY", "A"):
                for method in IO.METHODS:
                    got, _, _ = eng.db_similarity(R.encode(q), method)
                    assert same(got, IO.search(q, docs, method)), (q, method, bits)
        finally:
            eng.db_free()


def test_topk_is_the_stable_descending_sort(R, eng):
    rng = np.random.default_rng(32)
    docs = ["".join(rng.choice(list("AGCUN"), size=int(L))) for L in rng.integers(2, 40, size=6000)]
    eng.db_load(R.pack(docs, bits=4))
    try:
        q = docs[11]
        for method in ("set_jaccard_similarity", "multi_dice_similarity", "cosine", "pearson", "manhattan_distance"):
            alls, idx, sc = eng.db_similarity(R.encode(q), method, k=25)
            valid = np.where(~np.isnan(alls))[0]
            order = valid[np.argsort(-alls[valid], kind="stable")][:25]        # performance.py:13-14 on the non-NaN scores
            assert np.array_equal(idx, order), method
            assert np.array_equal(sc, alls[order]), method
    finally:
        eng.db_free()


def test_long_records_and_empty_query(R, eng):
    rng = np.random.default_rng(33)
    docs = ["".join(rng.choice(list(IO.SYMBOLS), size=2000)), "".join(rng.choice(list("AGCU"), size=1500)), "AG"]
    eng.db_load(R.pack(docs, bits=4))
    try:
        for method in IO.METHODS:
            got, _, _ = eng.db_similarity(R.encode(docs[0][:300]), method)
            assert same(got, IO.search(docs[0][:300], docs, method)), method
        got, _, _ = eng.db_similarity(np.zeros(0, np.uint8), "set_jaccard_similarity")
        assert np.array_equal(got, np.zeros(3))
        got, _, _ = eng.db_similarity(np.zeros(0, np.uint8), "cosine")
        assert np.isnan(got).all()
    finally:
        eng.db_free()


def test_dropin_search_collection_serves_every_measure(R, gir, dropin):
    """search_collection(query, 'tf', collection, IRMethods.<measure>) == the reference's list, and the measure
    objects called on two sequences give the same numbers (IR:443-477, IR:49-389)."""
    cwd = os.getcwd()
    os.chdir(dropin.cwd)
    try:
        IR = dropin.IR

        class Coll:
            def __init__(self, docs): self.docs = docs
            def find(self, flt): return iter(self.docs)
        coll = Coll([{"sequence": s} for s in gir["docs"]])
        q = gir["queries"][1]
        for name in IO.METHODS:
            res = IR.search_collection(q, 'tf', coll, getattr(IR, name))
            assert [s for s, _ in res] == gir["docs"]
            assert same(np.array([v for _, v in res]), unhex(gir["scores"][q][name])), name
        d = gir["docs"][3]
        want = {name: unhex(gir["scores"][q][name])[3] for name in IO.METHODS}
        assert IR.cosine(q, d) == want["cosine"] and IR.set_dice_similarity(q, d) == want["set_dice_similarity"]
        rd = {}
        IR.multi_jaccard_similarity(q, d, rd)
        assert rd == {"multi_jaccard_sim": want["multi_jaccard_similarity"]}
        got = {}
        IR.create_search_threads([IR.cosine, IR.set_jaccard_similarity, IR.wf_score], q, 'tf', coll,
                                 on_search_done=lambda r: got.__setitem__("mean", r), on_wf_done=lambda r: got.__setitem__("wf", r))
        assert len(got["wf"]) == len(gir["docs"]) and len(got["mean"]) == len(set(gir["docs"]))
    finally:
        os.chdir(cwd)


def test_sharded_similarity_topk_equals_single_shard(R, eng):
    """The similarity top-k over contiguous shards (local top-k + merge) == the top-k over the whole database."""
    from rna_sequence_diff_patch_b200.dist_search import shard_bounds, slice_packed
    from rna_sequence_diff_patch_b200.engine import topk_merge
    rng = np.random.default_rng(34)
    docs = ["".join(rng.choice(list("AGCUN"), size=int(L))) for L in rng.integers(2, 40, size=5000)]
    P = R.pack(docs, bits=4)
    q = R.encode(docs[7])
    for method in ("cosine", "set_dice_similarity", "multi_jaccard_similarity"):
        eng.db_load(P)
        try:
            _, wi, ws = eng.db_similarity(q, method, k=12, want_scores=False)
        finally:
            eng.db_free()
        parts_i, parts_s = [], []
        for lo, hi in shard_bounds(P.len, 3):
            eng.db_load(slice_packed(P, lo, hi), global_index_base=lo)
            try:
                _, i, s = eng.db_similarity(q, method, k=12, want_scores=False)
            finally:
                eng.db_free()
            parts_i.append(i[None, :]); parts_s.append(s[None, :])
        mi, ms = topk_merge(np.stack(parts_i), np.stack(parts_s))
        assert np.array_equal(mi[0], wi) and np.array_equal(ms[0], ws), method
