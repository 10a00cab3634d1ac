"""GPU parity under user-edited cost tables (the GUI lets the user type any float, gui.py:197,244):
integer, dyadic, non-dyadic, asymmetric, zero and 'substitution dearer than delete+insert' tables —
distances, scripts and search must stay bit-identical to the oracle in whatever mode gets chosen."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
SYM = "AGCUYRWSKMDVHBN"


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


def table(rng, kind):
    def val():
        if kind == "int": return float(rng.integers(0, 7))
        if kind == "dyadic": return float(rng.integers(0, 40)) / 8.0
        if kind == "float": return float(np.round(rng.random() * 3, 2))
        if kind == "dear": return float(rng.integers(3, 9))            # substitutions dearer than ins + del
        if kind == "big": return float(rng.integers(20, 90))           # does not fit the int8 tables
        raise ValueError(kind)
    ins, dele = (1.0, 1.0) if kind == "dear" else (val() + (0.0 if kind in ("int", "dyadic") else 0.01), val())
    if kind == "big":
        ins, dele = 70.0, 64.0
    upd = {a: {b: (0.0 if a == b else val()) for b in SYM} for a in SYM}
    return {"insert": ins, "delete": dele, "update": upd}


def rand_seqs(rng, n, lo, hi, alphabet):
    al = np.array(list(alphabet))
    return ["".join(al[rng.integers(0, len(al), size=L)]) for L in rng.integers(lo, hi + 1, size=n)]


@pytest.mark.parametrize("kind", ["int", "dyadic", "float", "dear", "big"])
def test_distance_script_search_under_random_tables(R, eng, kind):
    rng = np.random.default_rng(hash(kind) % 1000)
    for trial in range(3):
        costs = table(rng, kind)
        eng.set_costs(costs)
        for alphabet in ("AGCU", "AGCUN", SYM):
            a = rand_seqs(rng, 300, 1, 140, alphabet); b = rand_seqs(rng, 300, 1, 140, alphabet)
            ac, ao = O.concat(a); bc, bo = O.concat(b)
            want = O.distance_batch(ac, ao, bc, bo, costs)
            got = eng.distance_batch(R.pack(a), R.pack(b))
            assert np.array_equal(got, want), (kind, trial, alphabet, eng.last_mode)
            assert np.array_equal(eng.distance_batch(R.pack(a), R.pack(b), force_mode=3), want)
            # canonical scripts
            res = eng.script_batch(R.pack(a[:120]), R.pack(b[:120]), check_roundtrip=True)
            ops, oi, oj, cnt, dist = O.script_batch(*O.concat(a[:120]), *O.concat(b[:120]), costs)
            assert np.array_equal(res["dist"], dist) and np.array_equal(res["n_ops"], cnt) and res["ok"].all()
            for p in range(120):
                k = cnt[p]
                assert np.array_equal(res["op"][p, :k], ops[p, :k]), (kind, trial, alphabet, p, eng.last_mode)
                assert np.array_equal(res["oi"][p, :k], oi[p, :k])
        # search on short records (fast kernel when the table allows it, general path otherwise)
        recs = rand_seqs(rng, 4000, 20, 32, "AGCUN")
        qs = rand_seqs(rng, 3, 20, 32, "AGCUN")
        codes, off = O.concat(recs)
        eng.db_load(R.pack(recs, bits=4))
        idx, sc = eng.db_search_topk(R.pack(qs, bits=4), 8)
        eng.db_free()
        for q, query in enumerate(qs):
            wi, ws = O.search_topk(query, codes, off, costs, 8)
            assert np.array_equal(idx[q], wi) and np.array_equal(sc[q], ws), (kind, trial, q, eng.last_mode)


def test_modes_chosen_for_tables(R, eng):
    rng = np.random.default_rng(5)
    a = rand_seqs(rng, 64, 50, 100, "AGCU"); b = rand_seqs(rng, 64, 50, 100, "AGCU")
    A, B = R.pack(a), R.pack(b)
    for kind, mode in (("int", 1), ("dyadic", 1), ("dear", 1), ("big", 2)):
        costs = table(np.random.default_rng(1), kind)
        eng.set_costs(costs)
        eng.distance_batch(A, B)
        assert eng.last_mode == mode, (kind, eng.last_mode)
    costs = table(np.random.default_rng(1), "float")
    eng.set_costs(costs)
    eng.distance_batch(A, B)
    assert eng.last_mode in (2, 3)        # 2-decimal floats are dyadic only by accident


def test_zero_cost_insert_and_delete(R, eng, golden):
    costs = {"insert": 0.0, "delete": 0.0, "update": golden["default_costs"]["update"]}
    eng.set_costs(costs)
    rng = np.random.default_rng(9)
    a = rand_seqs(rng, 50, 1, 60, SYM); b = rand_seqs(rng, 50, 1, 60, SYM)
    got = eng.distance_batch(R.pack(a), R.pack(b))
    assert not got.any()
    res = eng.script_batch(R.pack(a), R.pack(b), check_roundtrip=True)
    ops, oi, oj, cnt, dist = O.script_batch(*O.concat(a), *O.concat(b), costs)
    assert np.array_equal(res["n_ops"], cnt) and res["ok"].all()
    for p in range(50):
        assert np.array_equal(res["op"][p, :cnt[p]], ops[p, :cnt[p]])
