"""GPU parity: canonical edit scripts (forward direction codes + traceback), packed -> reference
dict format, batched patch (prefix-sum scatter) and the on-device round-trip check."""
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


def rand_seqs(rng, n, lo, hi, alphabet):
    al = np.array(list(alphabet))
    return ["".join(al[rng.integers(0, len(al), size=L)]) for L in rng.integers(lo, hi + 1, size=n)]


def mutate(rng, s, alphabet, p_sub=0.05, p_ins=0.025, p_del=0.025):
    out = []
    for ch in s:
        r = rng.random()
        if r < p_sub: out.append(alphabet[rng.integers(len(alphabet))])
        elif r < p_sub + p_ins: out.extend([ch, alphabet[rng.integers(len(alphabet))]])
        elif r < p_sub + p_ins + p_del: pass
        else: out.append(ch)
    return "".join(out) or "A"


def check_against_oracle(R, eng, a, b, costs, force=0, expect_mode=None):
    eng.set_costs(costs)
    res = eng.script_batch(R.pack(a), R.pack(b), force_mode=force, check_roundtrip=True)
    if expect_mode is not None:
        assert eng.last_mode == expect_mode
    ac, ao = O.concat(a); bc, bo = O.concat(b)
    ops, oi, oj, cnt, dist = O.script_batch(ac, ao, bc, bo, costs)
    assert np.array_equal(res["dist"], dist)
    assert np.array_equal(res["n_ops"], cnt)
    for p in range(len(a)):
        k = cnt[p]
        assert np.array_equal(res["op"][p, :k], ops[p, :k]), (p, len(a[p]), len(b[p]))
        assert np.array_equal(res["oi"][p, :k], oi[p, :k]), p
        assert np.array_equal(res["oj"][p, :k], oj[p, :k]), p
    both = np.array([len(x) > 0 or len(y) > 0 for x, y in zip(a, b)])
    assert res["ok"].all()
    return res


def test_golden_scripts_are_paths0(R, eng, golden):
    from rna_sequence_diff_patch_b200 import sed
    cases = [c for c in [golden["G1"]] + golden["small"] + golden["medium"] + list(golden["xml_named"].values())
             if "es" in c]
    for user in (False, True):
        cs = [c for c in cases if c["user"] == user]
        costs = golden["user_costs" if user else "default_costs"]
        eng.set_costs(costs)
        for force in (0, 3):
            res = eng.script_batch(R.pack([c["a"] for c in cs], bits=4), R.pack([c["b"] for c in cs], bits=4),
                                   force_mode=force)
            for p, c in enumerate(cs):
                k = res["n_ops"][p]
                es = sed.es_from_packed(res["op"][p, :k], res["oi"][p, :k], res["oj"][p, :k], c["a"], c["b"])
                assert es == c["es"][0], (c["a"], c["b"], user, force)
                assert sed.format_edit_script(es) == c["es_fmt"][0]
                assert res["dist"][p] == c["distance"]
    assert len(cases) > 300


@pytest.mark.parametrize("table", ["default", "user"])
def test_acgu_scripts_vs_oracle(R, eng, golden, table):
    rng = np.random.default_rng(21)
    a = rand_seqs(rng, 600, 100, 300, "AGCU")
    b = [mutate(rng, s, "AGCU") if k % 10 else rand_seqs(rng, 1, 100, 300, "AGCU")[0] for k, s in enumerate(a)]
    costs = golden[f"{table}_costs"]
    check_against_oracle(R, eng, a, b, costs, expect_mode=2)
    check_against_oracle(R, eng, a, b, costs, force=3, expect_mode=3)


def test_iupac_scripts_fp64_vs_oracle(R, eng, golden):
    rng = np.random.default_rng(22)
    al = list(R.SYMBOLS)
    a = rand_seqs(rng, 400, 30, 200, al)
    b = [mutate(rng, s, al, 0.1, 0.05, 0.05) for s in a]
    for costs in (golden["default_costs"], golden["user_costs"]):
        check_against_oracle(R, eng, a, b, costs, expect_mode=3)


def test_long_pairs_multipass_and_edges(R, eng, golden):
    rng = np.random.default_rng(23)
    a = rand_seqs(rng, 48, 1000, 2000, "AGCU")
    b = [mutate(rng, s, "AGCU")[:2000] for s in a[:40]] + rand_seqs(rng, 8, 1000, 2000, "AGCU")
    check_against_oracle(R, eng, a, b, golden["default_costs"], expect_mode=2)
    check_against_oracle(R, eng, a[:12], b[:12], golden["default_costs"], force=3, expect_mode=3)
    lens = [0, 1, 2, 15, 16, 17, 31, 32, 33, 64, 65, 511, 513, 1025]
    al = np.array(list("AGCU"))
    aa, bb = [], []
    for la in lens:
        for lb in lens:
            aa.append("".join(al[rng.integers(0, 4, size=la)])); bb.append("".join(al[rng.integers(0, 4, size=lb)]))
    check_against_oracle(R, eng, aa, bb, golden["user_costs"])
    check_against_oracle(R, eng, aa, bb, golden["user_costs"], force=3)


def test_patch_batch_error_codes_and_tails(R, eng, golden):
    rng = np.random.default_rng(24)
    a = rand_seqs(rng, 300, 5, 120, "AGCU")
    b = [mutate(rng, s, "AGCU", 0.1, 0.05, 0.05) for s in a]
    costs = golden["default_costs"]
    eng.set_costs(costs)
    A, B = R.pack(a), R.pack(b)
    scripts = eng.script_batch(A, B)
    xs = []
    for k, s in enumerate(a):
        if k % 4 == 0: xs.append(s)                                           # exact source  -> code 0
        elif k % 4 == 1: xs.append(s + "".join(rng.choice(list("AGCU"), size=rng.integers(1, 6))))   # tail -> 1
        elif k % 4 == 2: xs.append(s[:-1] + ("G" if s[-1] != "G" else "A"))   # same length, differs  -> 1
        else: xs.append(s[:-1])                                               # shorter -> -1
    out, out_len, err = eng.patch_batch(scripts, A, B, R.pack(xs))
    for p in range(len(a)):
        k = scripts["n_ops"][p]
        code, want = O.patch_closed_codes(scripts["op"][p, :k], scripts["oi"][p, :k], scripts["oj"][p, :k],
                                          O.encode(a[p]), O.encode(b[p]),
                                          O.encode(xs[p]) if xs[p] else np.zeros(0, np.uint8))
        assert err[p] == code, p
        assert out_len[p] == len(want)
        assert np.array_equal(out[p, :out_len[p]], want)
        # and the reference's sequential semantics on the dict form agree (SED:380-457)
        from rna_sequence_diff_patch_b200 import sed
        es = sed.es_from_packed(scripts["op"][p, :k], scripts["oi"][p, :k], scripts["oj"][p, :k], a[p], b[p])
        assert sed.patching(es, xs[p]) == (code, O.decode(want))


def test_roundtrip_property_full_size_lengths(R, eng, golden):
    """C3-shaped pairs (1-2 kb, homologous + 10 % unrelated): on-device patch(script, A) == B."""
    rng = np.random.default_rng(20260003)
    a = rand_seqs(rng, 256, 1000, 2000, "AGCU")
    b = [(mutate(rng, s, "AGCU")[:2000] if k % 10 else rand_seqs(rng, 1, 1000, 2000, "AGCU")[0]) for k, s in enumerate(a)]
    eng.set_costs(golden["default_costs"])
    res = eng.script_batch(R.pack(a), R.pack(b), check_roundtrip=True)
    assert res["ok"].all()
    # distances agree with the distance-only kernel (different code path, int16x2)
    d = eng.distance_batch(R.pack(a), R.pack(b))
    assert np.array_equal(d, res["dist"])
    # script cost == distance: sum of op costs along the script
    for p in range(0, 256, 17):
        k = res["n_ops"][p]
        cost = 0.0
        for o, i, j in zip(res["op"][p, :k], res["oi"][p, :k], res["oj"][p, :k]):
            cost += 1.0 if o != 2 else (0.0 if a[p][i - 1] == b[p][j - 1] else 1.0)
        assert cost == res["dist"][p]


def test_edit_script_entry_point_any_length(R, eng, golden, monkeypatch):
    """sed.edit_script = generate_es(create_paths(wagnerFisher(a, b))[0], a, b) (SED:133-334) without the matrix: the
    reference's own scripts for short pairs (golden), the oracle's for a pair long enough for the panel-wavefront path,
    and the same long pair pushed through the row-block overflow path."""
    from rna_sequence_diff_patch_b200 import sed
    cases = [c for c in [golden["G1"]] + golden["small"][:40] + list(golden["xml_named"].values()) if "es" in c and c["a"] and c["b"]]
    for c in cases:
        costs = golden["user_costs" if c["user"] else "default_costs"]
        assert sed.edit_script(c["a"], c["b"], costs, engine=eng) == c["es"][0]
    with pytest.raises(IndexError):
        sed.edit_script("", "ACGU", golden["default_costs"], engine=eng)
    rng = np.random.default_rng(99)
    a = "".join(rng.choice(list("AGCU"), size=9000)); b = "".join(rng.choice(list("AGCU"), size=8700))
    costs = golden["default_costs"]
    ops, oi, oj, d = O.canonical_script(a, b, costs)
    want = sed.es_from_packed(ops, oi, oj, a, b)
    assert sed.edit_script(a, b, costs, engine=eng) == want
    assert sed.distance(a, b, costs, engine=eng) == d
    monkeypatch.setenv("RSD_LONG_BUDGET_MB", "6")
    assert sed.edit_script(a, b, costs, engine=eng) == want
    assert sed.patching(want, a) == (0, b)


def test_script_batch_routes_long_pairs_to_the_panel_kernels(R, eng, golden, monkeypatch):
    """A batch whose cells lie mostly in pairs of >= 4096 symbols runs on rsd_long_pairs (rings of panel pipelines)
    instead of one warp per pair: same packed scripts, distances and counts as the unrouted path and as the oracle;
    short pairs and an empty side in the same batch included."""
    rng = np.random.default_rng(4242)
    lens = [(5000, 5200), (4100, 300), (37, 4500), (6000, 6000), (800, 900), (0, 5), (4999, 4097)]
    a = ["".join(rng.choice(list("AGCU"), size=m)) for m, _ in lens]
    b = ["".join(rng.choice(list("AGCU"), size=n)) for _, n in lens]
    b[3] = a[3][:2500] + "".join(rng.choice(list("AGCU"), size=1000)) + a[3][3500:]          # homologous: long tie runs
    costs = golden["default_costs"]
    eng.set_costs(costs)
    A, B = R.pack(a), R.pack(b)
    launches0 = eng.launch_count()
    routed = eng.script_batch(A, B)
    n_routed = eng.launch_count() - launches0
    monkeypatch.setenv("RSD_SCRIPT_NO_LONG", "1")
    plain = eng.script_batch(A, B)
    assert n_routed <= 4                                  # one forward launch, traceback, emit
    for p in range(len(lens)):
        k = plain["n_ops"][p]
        assert routed["n_ops"][p] == k and routed["dist"][p] == plain["dist"][p]
        for key in ("op", "oi", "oj"):
            assert np.array_equal(routed[key][p, :k], plain[key][p, :k]), (p, key)
        if lens[p][0] and lens[p][1]:
            ops, oi, oj, d = O.canonical_script(a[p], b[p], costs)
            assert d == routed["dist"][p] and np.array_equal(routed["op"][p, :k], ops) and np.array_equal(routed["oj"][p, :k], oj)
