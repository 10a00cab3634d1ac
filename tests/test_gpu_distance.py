"""GPU parity: distances and matrices through the C ABI vs the oracle and the reference's golden
vectors.  Bit-exact (==) everywhere: integer modes are exact by construction, fp64 mode follows
the reference's operation order."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R():
    import __graft_entry__ as G
    G.build()
    import rna_sequence_diff_patch_b200 as R
    assert R.load_library().rsd_device_count() > 0, "no CUDA device: GPU tests need the B200 box"
    return R


@pytest.fixture(scope="module")
def eng(R):
    return R.Engine(0)


def rand_seqs(rng, n, lo, hi, alphabet):
    al = np.array(list(alphabet))
    lens = rng.integers(lo, hi + 1, size=n)
    return ["".join(al[rng.integers(0, len(al), size=L)]) for L in lens]


def oracle_batch(a, b, costs):
    ac, ao = O.concat(a); bc, bo = O.concat(b)
    return O.distance_batch(ac, ao, bc, bo, costs)


def test_dropin_golden_matrices(R, golden, dropin):
    cwd = os.getcwd()
    os.chdir(dropin.cwd)
    try:
        S = dropin.SED
        assert S.default_costs == golden["default_costs"] and S.user_costs == golden["user_costs"]
        dp = S.wagnerFisher('AGRGA', 'AGGGAA', True)
        assert str(dp) == golden["G1"]["dp_str"]
        paths = S.create_paths(dp)
        assert [S.generate_es(p, 'AGRGA', 'AGGGAA') for p in paths] == golden["G1"]["es"][:len(paths)]
        n = 0
        for c in [golden["G1"]] + golden["small"] + golden["medium"]:
            dp = S.wagnerFisher(c["a"], c["b"], c["user"])
            assert dp[len(dp) - 1][len(dp[0]) - 1].value == c["distance"]
            if "dp_repr" in c:
                assert [[repr(x.value) for x in row] for row in dp] == c["dp_repr"]
                assert dp.mask.tolist() == c["mask"]
            if "es" in c:
                paths = S.create_paths(dp)
                assert len(paths) == c["n_paths"]
                es0 = S.generate_es(paths[0], c["a"], c["b"])
                assert es0 == c["es"][0]
                assert list(S.patching(es0, c["a"])) == c["patch0"]
                rev = S.generate_rev_es(es0)
                assert list(S.patching(rev, c["b"])) == c["rev_patch0"]
            n += 1
        assert n > 300
        for c in golden["symbols"]:
            if "error" in c:
                with pytest.raises(KeyError):
                    S.wagnerFisher(c["a"], c["b"])
            else:
                dp = S.wagnerFisher(c["a"], c["b"])
                assert [[repr(x.value) for x in row] for row in dp] == c["dp_repr"]
        for c in golden["empties"]:
            dp = S.wagnerFisher(c["a"], c["b"])
            assert [[repr(x.value) for x in row] for row in dp] == c["dp_repr"]
    finally:
        os.chdir(cwd)


def test_xml_600_pairs_both_tables(R, eng, golden):
    seqs = golden["xml_seqs"]
    ii = [r[0] for r in golden["xml_all_pairs"]]; jj = [r[1] for r in golden["xml_all_pairs"]]
    A = R.pack([seqs[i] for i in ii]); B = R.pack([seqs[j] for j in jj])
    for col, costs in ((2, golden["default_costs"]), (3, golden["user_costs"])):
        eng.set_costs(costs)
        want = np.array([r[col] for r in golden["xml_all_pairs"]])
        for force in (0, 2, 3):
            got = eng.distance_batch(A, B, force_mode=force)
            assert np.array_equal(got, want), (col, force)
    assert eng.last_mode == 3


def test_g6_fp64_fingerprints(R, eng, golden):
    eng.set_costs(golden["default_costs"])
    A = R.pack([c["a"] for c in golden["G6"]]); B = R.pack([c["b"] for c in golden["G6"]])
    got = eng.distance_batch(A, B)
    assert eng.last_mode == 3
    assert got.tolist() == [c["distance"] for c in golden["G6"]]
    assert got.tolist() == [66.74999999999994, 125.39999999999989, 191.40999999999983, 315.9600000000003,
                            628.7499999999999]


@pytest.mark.parametrize("table", ["user", "default"])
def test_acgu_batch_all_modes_agree_with_oracle(R, eng, golden, table):
    costs = golden[f"{table}_costs"]
    eng.set_costs(costs)
    rng = np.random.default_rng(11)
    a = rand_seqs(rng, 3000, 100, 300, "AGCU"); b = rand_seqs(rng, 3000, 100, 300, "AGCU")
    want = oracle_batch(a, b, costs)
    A, B = R.pack(a), R.pack(b)
    assert A.bits == 2
    got = eng.distance_batch(A, B)
    assert eng.last_mode == 1, "ACGU with integer costs must take the int16x2 DPX path"
    assert np.array_equal(got, want)
    for force in (2, 3):
        assert np.array_equal(eng.distance_batch(A, B, force_mode=force), want), force
    A4, B4 = A.repack(4), B.repack(4)
    for force in (0, 2, 3):
        assert np.array_equal(eng.distance_batch(A4, B4, force_mode=force), want), force


def test_iupac_batch_fp64(R, eng, golden):
    rng = np.random.default_rng(12)
    a = rand_seqs(rng, 2000, 100, 300, R.SYMBOLS); b = rand_seqs(rng, 2000, 100, 300, R.SYMBOLS)
    for costs in (golden["default_costs"], golden["user_costs"]):
        eng.set_costs(costs)
        got = eng.distance_batch(R.pack(a), R.pack(b))
        assert eng.last_mode == 3
        assert np.array_equal(got, oracle_batch(a, b, costs))


def test_dyadic_subset_int32_scaled(R, eng, golden):
    """{A,G,C,U,Y,R,W,S,K,M,N} under default costs: everything is a multiple of 0.25 -> int32 x4."""
    rng = np.random.default_rng(13)
    al = "AGCUYRWSKMN"
    a = rand_seqs(rng, 1500, 20, 120, al); b = rand_seqs(rng, 1500, 20, 120, al)
    eng.set_costs(golden["default_costs"])
    got = eng.distance_batch(R.pack(a), R.pack(b))
    assert eng.last_mode == 2
    want = oracle_batch(a, b, golden["default_costs"])
    assert np.array_equal(got, want)
    assert np.array_equal(eng.distance_batch(R.pack(a), R.pack(b), force_mode=3), want)


def test_edge_lengths_and_multipass(R, eng, golden):
    """Empty and ragged inputs, strip boundaries (C = 32 / 16 columns per lane), > 32 strips."""
    rng = np.random.default_rng(14)
    lens = [0, 1, 2, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 96, 97, 511, 512, 513, 1023, 1024, 1025, 1100, 2050]
    al = np.array(list("AGCU"))
    a, b = [], []
    for la in lens:
        for lb in (0, 1, 31, 32, 33, 64, 100, 513, 1024, 1025, 2049):
            if la * lb > 1_300_000:
                continue
            a.append("".join(al[rng.integers(0, 4, size=la)])); b.append("".join(al[rng.integers(0, 4, size=lb)]))
    for costs in (golden["user_costs"], golden["default_costs"]):
        eng.set_costs(costs)
        want = oracle_batch(a, b, costs)
        A, B = R.pack(a), R.pack(b)
        for force in (0, 2, 3):
            got = eng.distance_batch(A, B, force_mode=force)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (force, [(len(a[k]), len(b[k]), got[k], want[k]) for k in bad[:5]])
    # the same shapes over the full alphabet (fp64, 4-bit)
    al15 = np.array(list(R.SYMBOLS))
    a15 = ["".join(al15[rng.integers(0, 15, size=len(x))]) for x in a]
    b15 = ["".join(al15[rng.integers(0, 15, size=len(x))]) for x in b]
    eng.set_costs(golden["default_costs"])
    got = eng.distance_batch(R.pack(a15), R.pack(b15))
    assert np.array_equal(got, oracle_batch(a15, b15, golden["default_costs"]))


def test_single_pair_and_tiny_batches(R, eng, golden):
    eng.set_costs(golden["default_costs"])
    for n in (1, 2, 3, 33):
        a = ["ACGUACGUAC"] * n; b = ["ACGUUCGUA"] * n
        got = eng.distance_batch(R.pack(a), R.pack(b))
        assert got.tolist() == [O.distance(a[0], b[0], golden["default_costs"])] * n
    assert eng.distance_batch(R.pack([]), R.pack([])).shape == (0,)


def test_large_batch_property_identity_and_symmetry(R, eng, golden):
    """BASELINE-size property checks that need no oracle: d(x,x) == 0; with ins == del and a
    symmetric table d(a,b) == d(b,a); duplicating a batch duplicates its answers."""
    eng.set_costs(golden["default_costs"])
    rng = np.random.default_rng(15)
    a = rand_seqs(rng, 60000, 100, 300, "AGCU"); b = rand_seqs(rng, 60000, 100, 300, "AGCU")
    A, B = R.pack(a), R.pack(b)
    d_ab = eng.distance_batch(A, B); d_ba = eng.distance_batch(B, A)
    assert np.array_equal(d_ab, d_ba)
    assert not eng.distance_batch(A, A).any()
    sub = rng.choice(60000, size=400, replace=False)
    want = oracle_batch([a[k] for k in sub], [b[k] for k in sub], golden["default_costs"])
    assert np.array_equal(d_ab[sub], want)


def test_wf_score_and_search_collection_g7(R, golden, dropin):
    cwd = os.getcwd()
    os.chdir(dropin.cwd)
    try:
        IR = dropin.IR

        class Coll:
            def __init__(self, docs): self.docs = docs
            def find(self, flt): return iter(self.docs)
        coll = Coll([{"sequence": s} for s in golden["xml_seqs"]])
        for c in golden["G7"]:
            scores = IR.search_collection(c["query"], 'tf', coll, IR.wf_score)
            assert [list(x) for x in scores] == c["scores"]
            assert [list(x) for x in IR.top_k(scores, 6)] == c["top6"]
        for a, b, v in golden["wf_score_user"]:
            assert IR.wf_score(a, b, True) == v
        # create_search_threads replacement (IR:480-515): one GPU scan, both callbacks, packed-DB collection
        from rna_sequence_diff_patch_b200.ingest import SequenceDB
        db = SequenceDB(dict(zip(golden["xml_ids"], golden["xml_seqs"])))
        got = {}
        IR.create_search_threads([IR.wf_score], golden["G7"][0]["query"], 'tf', db,
                                 on_search_done=lambda r: got.__setitem__("mean", r),
                                 on_wf_done=lambda r: got.__setitem__("wf", r))
        assert [list(x) for x in got["wf"]] == golden["G7"][0]["scores"]
        assert [list(x) for x in IR.top_k(got["mean"], 6)] == golden["G7"][0]["top6"]
    finally:
        os.chdir(cwd)


def test_shared_sequence_layout_falls_back_to_whole_copy(R, eng, golden):
    """One query shared by every pair (all a_start equal): the chunked host copy must not be used."""
    import ctypes as C
    from rna_sequence_diff_patch_b200 import _lib
    rng = np.random.default_rng(16)
    n = 70000                                            # >= 65536 pairs -> chunking would kick in
    b = rand_seqs(rng, n, 20, 40, "AGCU")
    q = "ACGUACGGUUACGCAUUCGA"
    B = R.pack(b); Q = R.pack([q])
    a_start = np.zeros(n, np.int64); a_len = np.full(n, len(q), np.int32)
    out = np.zeros(n)
    eng.set_costs(golden["default_costs"])
    mode = C.c_int()
    _lib.check(R.load_library().rsd_distance_batch(
        eng.ctx, _lib.ptr(Q.words, C.c_uint32), _lib.ptr(a_start, C.c_int64), _lib.ptr(a_len, C.c_int32), Q.words.shape[0],
        _lib.ptr(B.words, C.c_uint32), _lib.ptr(B.start, C.c_int64), _lib.ptr(B.len, C.c_int32), B.words.shape[0],
        n, len(q), 40, 2, 0xF, 0, _lib.ptr(out, C.c_double), C.byref(mode)))
    want = oracle_batch([q] * n, b, golden["default_costs"])
    assert np.array_equal(out, want)
    # sequences stored in REVERSE pair order: start[] is monotone at no chunk boundary either
    order = np.arange(n)[::-1]
    Br = R.pack([b[k] for k in order])
    b_start = Br.start[order].copy(); b_len = Br.len[order].copy()
    A = R.pack([q] * n)
    out2 = np.zeros(n)
    _lib.check(R.load_library().rsd_distance_batch(
        eng.ctx, _lib.ptr(A.words, C.c_uint32), None, _lib.ptr(A.len, C.c_int32), A.words.shape[0],
        _lib.ptr(Br.words, C.c_uint32), _lib.ptr(b_start, C.c_int64), _lib.ptr(b_len, C.c_int32), Br.words.shape[0],
        n, len(q), 40, 2, 0xF, 0, _lib.ptr(out2, C.c_double), C.byref(mode)))
    assert np.array_equal(out2, want)
    # in pair order for the first chunks, out of order only from pair 50 000 on (the tail stored reversed): the chunks
    # already sent stay valid, the whole buffer follows for the rest
    order3 = np.concatenate([np.arange(50000), np.arange(50000, n)[::-1]])
    Bt = R.pack([b[k] for k in order3])
    inv = np.empty(n, np.int64); inv[order3] = np.arange(n)
    b_start3 = Bt.start[inv].copy(); b_len3 = Bt.len[inv].copy()
    out3 = np.zeros(n)
    _lib.check(R.load_library().rsd_distance_batch(
        eng.ctx, _lib.ptr(A.words, C.c_uint32), None, _lib.ptr(A.len, C.c_int32), A.words.shape[0],
        _lib.ptr(Bt.words, C.c_uint32), _lib.ptr(b_start3, C.c_int64), _lib.ptr(b_len3, C.c_int32), Bt.words.shape[0],
        n, len(q), 40, 2, 0xF, 0, _lib.ptr(out3, C.c_double), C.byref(mode)))
    assert np.array_equal(out3, want)


def test_canonical_layout_without_start_equals_explicit_start(R, eng, golden):
    """start == NULL (rsd_pack's layout, offsets rebuilt on the device) vs caller-supplied start[]: same distances;
    lengths 0 and exact word multiples included, batch large enough for the chunked copies."""
    import ctypes as C
    from rna_sequence_diff_patch_b200 import _lib
    rng = np.random.default_rng(18)
    n = 150_001
    la = rng.integers(0, 70, size=n); lb = rng.integers(0, 70, size=n)
    la[rng.random(n) < 0.05] = 0; lb[rng.random(n) < 0.05] = 32; la[:3] = [16, 0, 0]; lb[-2:] = [0, 0]
    oa = np.zeros(n + 1, np.int64); np.cumsum(la, out=oa[1:]); ob = np.zeros(n + 1, np.int64); np.cumsum(lb, out=ob[1:])
    for alpha, table in ((4, "user_costs"), (15, "default_costs")):
        ca = rng.integers(0, alpha, size=int(oa[-1]), dtype=np.uint8); cb = rng.integers(0, alpha, size=int(ob[-1]), dtype=np.uint8)
        costs = golden[table]
        eng.set_costs(costs)
        A, B = R.pack((ca, oa)), R.pack((cb, ob))
        assert A.canonical and B.canonical
        got = eng.distance_batch(A, B)                                  # start == NULL on both sides
        want = O.distance_batch(ca, oa, cb, ob, costs)
        assert np.array_equal(got, want)
        A.canonical = False                                             # explicit start on one side, NULL on the other
        assert np.array_equal(eng.distance_batch(A, B), want)
        # lengths that need more words than the buffer holds are refused
        mode = C.c_int(); out = np.zeros(n)
        rc = R.load_library().rsd_distance_batch(
            eng.ctx, _lib.ptr(A.words, C.c_uint32), None, _lib.ptr(A.len, C.c_int32), A.words.shape[0] - 40,
            _lib.ptr(B.words, C.c_uint32), None, _lib.ptr(B.len, C.c_int32), B.words.shape[0],
            n, 0, 0, A.bits, A.symmask | B.symmask, 0, _lib.ptr(out, C.c_double), C.byref(mode))
        assert rc == _lib.RSD_EINVAL


@pytest.mark.parametrize("n", [1, 4097, 200_003])
def test_distance_from_raw_codes_packs_on_the_device(R, eng, golden, n):
    """rsd_distance_batch_codes: 1 byte per symbol in, packed by k_pack_codes chunk by chunk == the oracle."""
    rng = np.random.default_rng(19 + n)
    la = rng.integers(0, 90, size=n).astype(np.int32); lb = rng.integers(0, 90, size=n).astype(np.int32)
    la[rng.random(n) < 0.03] = 0; lb[rng.random(n) < 0.03] = 0
    oa = np.zeros(n + 1, np.int64); np.cumsum(la, out=oa[1:]); ob = np.zeros(n + 1, np.int64); np.cumsum(lb, out=ob[1:])
    for alpha, bits, table, mode in ((4, 2, "user_costs", 1), (4, 4, "default_costs", 2), (15, 4, "default_costs", 3)):
        ca = rng.integers(0, alpha, size=int(oa[-1]), dtype=np.uint8); cb = rng.integers(0, alpha, size=int(ob[-1]), dtype=np.uint8)
        costs = golden[table]
        eng.set_costs(costs)
        got = eng.distance_batch_codes(ca, la, cb, lb, bits=bits, symmask=0xF if alpha == 4 else 0)
        assert eng.last_mode == mode
        assert np.array_equal(got, O.distance_batch(ca, oa, cb, ob, costs))
    if n > 1000:
        bad = ca.copy(); bad[int(oa[n // 2]) + 1 if la[n // 2] > 1 else 0] = 9
        with pytest.raises(R.RsdError):
            eng.distance_batch_codes(bad % 16, la, cb % 4, lb, bits=2)


def test_plan_slots_start_clean_on_recycled_device_memory(R, golden):
    """The plan's privatised bin counters must be zeroed by the library, not by luck: fill a large part of the
    device with 0xFF, release it, then let a NEW context allocate its plan slots (chunked call: slots 0..4)."""
    import torch
    junk = [torch.full((1 << 28,), -1, dtype=torch.int32, device="cuda") for _ in range(8)]       # 8 GiB of 0xFF
    torch.cuda.synchronize()
    del junk
    torch.cuda.empty_cache()
    eng2 = R.Engine(0)
    try:
        rng = np.random.default_rng(20)
        a = rand_seqs(rng, 80000, 30, 90, "AGCU"); b = rand_seqs(rng, 80000, 30, 90, "AGCU")
        eng2.set_costs(golden["user_costs"])
        got = eng2.distance_batch(R.pack(a), R.pack(b))
        assert np.array_equal(got, oracle_batch(a, b, golden["user_costs"]))
        res = eng2.script_batch(R.pack(a[:3000]), R.pack(b[:3000]))
        ac, ao = O.concat(a[:3000]); bc, bo = O.concat(b[:3000])
        _, _, _, cnt, dist = O.script_batch(ac, ao, bc, bo, golden["user_costs"])
        assert np.array_equal(res["n_ops"], cnt) and np.array_equal(res["dist"], dist)
    finally:
        eng2.close()


def test_pair_list_shards_concatenate_to_the_unsharded_result(R, eng, golden):
    """8e: contiguous pair ranges balanced by cells, scored independently (3 emulated ranks on one GPU)."""
    from rna_sequence_diff_patch_b200.dist_pairs import ShardedPairs
    rng = np.random.default_rng(17)
    a = rand_seqs(rng, 4000, 0, 90, "AGCUN"); b = rand_seqs(rng, 4000, 0, 90, "AGCUN")
    A = R.pack(a); B = R.pack(b)
    eng.set_costs(golden["default_costs"])
    whole = eng.distance_batch(A, B)
    parts = []
    for rank in range(3):
        sp = ShardedPairs(eng, rank, 3)
        bounds, (lo, hi) = sp.local_range(A, B)
        parts.append(sp.distance_batch(A, B))                 # no process group: returns the local range
        assert parts[-1].shape[0] == hi - lo
    assert np.array_equal(np.concatenate(parts), whole)
    assert np.array_equal(whole, oracle_batch(a, b, golden["default_costs"]))


def test_distance_batch_routes_long_pairs_to_the_panel_kernels(R, eng, golden, monkeypatch):
    """Batches whose cells lie mostly in pairs of >= 4096 symbols run on rsd_long_pairs (distance only): same numbers as
    the tape-pass kernels and the oracle, from packed words (explicit and canonical layout) and from raw codes."""
    import ctypes as C
    from rna_sequence_diff_patch_b200 import _lib
    rng = np.random.default_rng(777)
    lens = [(5000, 4800), (4200, 100), (6000, 6100), (300, 350), (0, 7), (4096, 4096)]
    a = ["".join(rng.choice(list("AGCU"), size=m)) for m, _ in lens]
    b = ["".join(rng.choice(list("AGCU"), size=n)) for _, n in lens]
    for costs in (golden["default_costs"], golden["user_costs"]):
        eng.set_costs(costs)
        A, B = R.pack(a), R.pack(b)
        routed = eng.distance_batch(A, B)
        monkeypatch.setenv("RSD_DIST_NO_LONG", "1")
        plain = eng.distance_batch(A, B)
        monkeypatch.delenv("RSD_DIST_NO_LONG")
        want = oracle_batch(a, b, costs)
        assert np.array_equal(routed, want) and np.array_equal(plain, want)
        # raw codes (1 byte per symbol) through rsd_distance_batch_codes
        ca = np.concatenate([O.encode(x) for x in a]); cb = np.concatenate([O.encode(x) for x in b])
        la = np.array([len(x) for x in a], np.int32); lb = np.array([len(x) for x in b], np.int32)
        out = np.zeros(len(a)); mode = C.c_int()
        _lib.check(R.load_library().rsd_distance_batch_codes(
            eng.ctx, _lib.ptr(ca, C.c_uint8), _lib.ptr(la, C.c_int32), _lib.ptr(cb, C.c_uint8), _lib.ptr(lb, C.c_int32),
            len(a), 0, 0, 2, 0xF, 0, _lib.ptr(out, C.c_double), C.byref(mode)))
        assert np.array_equal(out, want)
