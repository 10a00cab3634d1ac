"""GPU parity: the long-pair path (panel wavefront + traceback) vs the oracle and via properties."""
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


def mutate_codes(rng, a, alpha=4, p_sub=0.05, p_ins=0.025, p_del=0.025):
    r = rng.random(a.shape[0])
    keep = r >= p_del
    out = []
    sub = rng.integers(0, alpha, size=a.shape[0], dtype=np.uint8)
    ins = rng.integers(0, alpha, size=a.shape[0], dtype=np.uint8)
    for k in range(a.shape[0]):
        if not keep[k]:
            continue
        out.append(sub[k] if r[k] < p_del + p_sub else a[k])
        if r[k] > 1 - p_ins:
            out.append(ins[k])
    return np.array(out, dtype=np.uint8)


def apply_script(op, oj, b):
    """closed-form patch on the host: destination symbols of the non-delete ops"""
    keep = op != 1
    return b[oj[keep] - 1]


@pytest.mark.parametrize("shape", [(3000, 3000), (5000, 777), (300, 4100), (1, 700), (257, 256), (2049, 2048)])
def test_long_pair_vs_oracle(R, eng, golden, shape):
    rng = np.random.default_rng(sum(shape))
    m, n = shape
    a = rng.integers(0, 4, size=m, dtype=np.uint8)
    b = mutate_codes(rng, a)[:n] if m == n else rng.integers(0, 4, size=n, dtype=np.uint8)
    for costs, modes in ((golden["default_costs"], (0, 3)), (golden["user_costs"], (0,))):
        eng.set_costs(costs)
        ops, oi, oj, d = O.canonical_script(O.decode(a), O.decode(b), costs)
        for force in modes:
            res = eng.long_pair(a, b, force_mode=force)
            assert res["mode"] == (3 if force == 3 else 2)
            assert res["dist"] == d
            assert np.array_equal(res["op"], ops) and np.array_equal(res["oi"], oi) and np.array_equal(res["oj"], oj)


def test_long_pair_iupac_fp64(R, eng, golden):
    rng = np.random.default_rng(77)
    a = rng.integers(0, 15, size=2500, dtype=np.uint8)
    b = mutate_codes(rng, a, alpha=15, p_sub=0.1)
    eng.set_costs(golden["default_costs"])
    res = eng.long_pair(a, b)
    assert res["mode"] == 3
    ops, oi, oj, d = O.canonical_script(O.decode(a), O.decode(b), golden["default_costs"])
    assert res["dist"] == d and np.array_equal(res["op"], ops) and np.array_equal(res["oi"], oi)


def test_empty_sides(R, eng, golden):
    eng.set_costs(golden["user_costs"])
    a = np.array([0, 1, 2], np.uint8); e = np.zeros(0, np.uint8)
    r = eng.long_pair(a, e)
    assert r["dist"] == 9.0 and r["op"].tolist() == [1, 1, 1] and r["oi"].tolist() == [1, 2, 3]
    r = eng.long_pair(e, a)
    assert r["dist"] == 6.0 and r["op"].tolist() == [0, 0, 0] and r["oj"].tolist() == [1, 2, 3]


def test_c4_shape_properties(R, eng, golden):
    """50 kb x 50 kb (BASELINE config 4): no CPU oracle at this size in a test budget; check
    (i) the script is a valid path whose cost equals the reported distance, (ii) patching A with it
    gives B, (iii) the distance agrees with the batched distance kernel (independent code path)."""
    rng = np.random.default_rng(20260004)
    a = rng.integers(0, 4, size=50000, dtype=np.uint8)
    b = mutate_codes(rng, a)
    eng.set_costs(golden["default_costs"])
    res = eng.long_pair(a, b)
    op, oi, oj = res["op"], res["oi"], res["oj"]
    assert oi[-1] == a.shape[0] and oj[-1] == b.shape[0]
    di = np.diff(np.concatenate([[0], oi])); dj = np.diff(np.concatenate([[0], oj]))
    assert np.array_equal(di, (op != 0).astype(np.int64)) and np.array_equal(dj, (op != 1).astype(np.int64))
    upd = op == 2
    cost = float((op != 2).sum() + (a[oi[upd] - 1] != b[oj[upd] - 1]).sum())
    assert cost == res["dist"]
    assert np.array_equal(apply_script(op, oj, b), b)
    off_a = np.array([0, a.shape[0]], np.int64); off_b = np.array([0, b.shape[0]], np.int64)
    d = eng.distance_batch(R.pack((a, off_a)), R.pack((b, off_b)))
    assert d[0] == res["dist"]


def test_wide_key_fallback_paths_agree(R, eng, golden, monkeypatch):
    """The 32-bit modular-key kernel is the default; the double-carried integer key kernel stays as the
    fallback (forced here through RSD_LONG_WIDE, and reached naturally by costs too large for the 32-bit bound)."""
    rng = np.random.default_rng(77)
    a = rng.integers(0, 4, size=2600, dtype=np.uint8)
    b = mutate_codes(rng, a)
    costs = golden["default_costs"]
    eng.set_costs(costs)
    ops, oi, oj, d = O.canonical_script(O.decode(a), O.decode(b), costs)
    fast = eng.long_pair(a, b)
    monkeypatch.setenv("RSD_LONG_WIDE", "1")
    wide = eng.long_pair(a, b)
    monkeypatch.delenv("RSD_LONG_WIDE")
    monkeypatch.setenv("RSD_LONG_R1", "1")                      # 32-bit keys, one row per step
    one_row = eng.long_pair(a, b)
    monkeypatch.delenv("RSD_LONG_R1")
    for res in (fast, wide, one_row):
        assert res["mode"] == 2 and res["dist"] == d
        assert np.array_equal(res["op"], ops) and np.array_equal(res["oi"], oi) and np.array_equal(res["oj"], oj)
    # integer costs in the thousands: (32 * maxc + 64) << S exceeds 2^30 -> the wide kernel is chosen by the library
    big = {"insert": 3000.0, "delete": 2500.0,
           "update": {x: {y: (0.0 if x == y else 4100.0 + 7 * ((ord(x) + 3 * ord(y)) % 11)) for y in "AGCUYRWSKMDVHBN"} for x in "AGCUYRWSKMDVHBN"}}
    eng.set_costs(big)
    ops, oi, oj, d = O.canonical_script(O.decode(a), O.decode(b), big)
    res = eng.long_pair(a, b)
    assert res["mode"] == 2 and res["dist"] == d
    assert np.array_equal(res["op"], ops) and np.array_equal(res["oi"], oi) and np.array_equal(res["oj"], oj)
