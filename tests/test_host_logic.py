"""Host-side mirror of StringEditDistance.py (sed.py) against the reference's golden vectors.
The device-computed inputs (values + tie masks) are replaced here by the oracle's, so this runs
without a GPU; tests/test_gpu_*.py run the same checks end to end through librsd.so."""
import numpy as np
import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def sed():
    import __graft_entry__ as G
    G.build()
    from rna_sequence_diff_patch_b200 import sed
    return sed


def costs_of(g, user):
    return g["user_costs"] if user else g["default_costs"]


def dp_from_oracle(sed, c, costs):
    D, M = O.matrix(c["a"], c["b"], costs)
    return sed.DPMatrix(D, M, c["a"], c["b"], costs)


def cases(g):
    return [g["G1"]] + g["small"] + g["medium"] + list(g["xml_named"].values())


def test_paths_order_es_and_format(sed, golden):
    n = 0
    for c in cases(golden):
        if "paths_cells" not in c:
            continue
        costs = costs_of(golden, c["user"])
        dp = dp_from_oracle(sed, c, costs)
        keep = len(c["paths_cells"])
        paths = sed.create_paths(dp, limit=None if c["n_paths"] <= 4000 else keep)
        if c["n_paths"] <= 4000:
            assert len(paths) == c["n_paths"]
        got = [[[nd.i + 1, nd.j + 1] for nd in p] for p in paths[:keep]]
        assert got == c["paths_cells"]
        ess = [sed.generate_es(p, c["a"], c["b"]) for p in paths[:keep]]
        assert ess[:3] == c["es"]
        assert [sed.format_edit_script(e) for e in ess] == c["es_fmt"]
        assert sed.canonical_path(dp) == paths[0]
        es0 = ess[0]
        assert list(sed.patching(es0, c["a"])) == c["patch0"]
        rev = sed.generate_rev_es(es0)
        assert rev == c["rev0"]
        assert list(sed.patching(rev, c["b"])) == c["rev_patch0"]
        assert sed.generate_sequence_from_es(es0) == c["seq_from_es0"]
        n += 1
    assert n > 250


def test_dp_view_repr_and_typing(sed, golden):
    for c in [golden["G1"]] + golden["small"][:100] + golden["medium"][:20]:
        if "dp_repr" not in c:
            continue
        dp = dp_from_oracle(sed, c, costs_of(golden, c["user"]))
        assert len(dp) == len(c["a"]) + 1 and len(dp[0]) == len(c["b"]) + 1
        assert [[repr(x.value) for x in row] for row in dp] == c["dp_repr"]
        assert dp[len(dp) - 1][len(dp[0]) - 1].value == c["distance"]
    g1 = dp_from_oracle(sed, golden["G1"], golden["user_costs"])
    assert str(g1) == golden["G1"]["dp_str"]
    assert golden["import_stdout"].startswith(str(g1))


def test_int_costs_typing(sed):
    """JSON integer literals propagate Python ints further than the diagonal."""
    costs = {"insert": 1, "delete": 2.0, "update": {a: {b: (1 if a < b else 1.5) for b in "AGCU"} for a in "AGCU"}}
    for a, b in [("AGCU", "GCUA"), ("AAGG", "AGGG"), ("ACGU", "ACGU")]:
        want, _ = O.py_matrix(a, b, costs)
        oc = {"insert": 1.0, "delete": 2.0, "update": {x: {y: float(v) for y, v in r.items()} for x, r in costs["update"].items()}}
        for x in "YRWSKMDVHBN":
            oc["update"][x] = {}
        D, M = O.matrix(a, b, oc)
        dp = sed.DPMatrix(D, M, a, b, costs)
        assert [[repr(v.value) for v in row] for row in dp] == [[repr(v) for v in row] for row in want]


def test_patch_error_codes_and_hand_edited_scripts(sed, golden):
    es1 = golden["G1"]["es"][0]
    for x, out in golden["G1"]["patch_cases"]:
        assert list(sed.patching(es1, x)) == out
    for c in golden["patch_odd"]:
        assert list(sed.patching(c["es"], c["x"])) == c["out"]


def test_node_edges_match_reference_structure(sed, golden):
    c = golden["G1"]
    dp = dp_from_oracle(sed, c, golden["user_costs"])
    m, n = dp.m, dp.n
    for r in range(m + 1):
        for col in range(n + 1):
            nd = dp[r][col]
            assert (nd.i, nd.j) == (r - 1, col - 1)
            ops = [e.operation for e in nd.incoming_edges]
            want = [o for k, o in enumerate(("insert", "delete", "update")) if c["mask"][r][col] >> k & 1]
            assert ops == want
            for e in nd.edges:
                assert e.source == nd and any(x.source == nd for x in e.destination.incoming_edges)
    assert [e.operation for e in dp[1][0].edges][0] == "delete"      # column-0 order (SED:167-182)


def test_symbol_errors_like_reference(sed, golden):
    for c in golden["symbols"]:
        if "error" in c:
            with pytest.raises(KeyError) as ei:
                sed._validate_and_encode(c["a"], c["b"], golden["default_costs"])
            assert ei.value.args[0] == c["key"]
        else:
            ca, cb = sed._validate_and_encode(c["a"], c["b"], golden["default_costs"])
            assert len(ca) == len(c["a"]) and len(cb) == len(c["b"])


def test_empty_strings(sed, golden):
    for c in golden["empties"]:
        D, M = O.matrix(c["a"], c["b"], golden["default_costs"])
        dp = sed.DPMatrix(D, M, c["a"], c["b"], golden["default_costs"])
        assert [[repr(x.value) for x in row] for row in dp] == c["dp_repr"]
        paths = sed.create_paths(dp)
        assert [[[nd.i + 1, nd.j + 1] for nd in p] for p in paths] == [[list(x) for x in p] for p in c["paths_cells"]]
        if c.get("es_error") == "IndexError":
            with pytest.raises(IndexError):
                sed.generate_es(paths[0], c["a"], c["b"])
        else:
            assert sed.generate_es(paths[0], c["a"], c["b"]) == c["es"]


def test_cost_and_min_cost_helpers(sed, golden):
    dc = golden["default_costs"]
    assert sed.cost("a", "A", dc) == 0 and isinstance(sed.cost("A", "A", dc), int)
    assert sed.cost("A", "R", dc) == 0.5
    with pytest.raises(KeyError):
        sed.cost("a", "G", dc)
    c = golden["G1"]
    dp = dp_from_oracle(sed, c, golden["user_costs"])
    val, ops = sed.min_cost(dp, 3, 3, c["a"], c["b"], golden["user_costs"])
    assert val == 0.5 and ops == [None, None, (2, 2, "update")]


def test_hang_witness_terminates(sed, golden):
    """'AAA'->'GGCUUGU' deadlocks the reference's bounded queue; ours returns."""
    D, M = O.matrix("AAA", "GGCUUGU", golden["default_costs"])
    dp = sed.DPMatrix(D, M, "AAA", "GGCUUGU", golden["default_costs"])
    paths = sed.create_paths(dp)
    want = O.all_paths(M.tolist())
    assert [[(nd.i + 1, nd.j + 1) for nd in p] for p in paths] == want
