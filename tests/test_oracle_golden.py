"""Pins oracle/ (C restatement + Python restatements) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O


def costs_of(g, user):
    return g["user_costs"] if user else g["default_costs"]


def all_cases(g):
    return [g["G1"]] + g["small"] + g["medium"] + list(g["xml_named"].values())


def test_builtin_cost_tables_match_reference_files(golden, tmp_path):
    """cost_tables.py == the dicts the reference loaded from its costs.json / user_costs.json (values, key order,
    float typing), and write_cost_files() round-trips through the JSON format."""
    import json, os
    from rna_sequence_diff_patch_b200 import cost_tables as T
    assert json.dumps(T.DEFAULT_COSTS) == json.dumps(golden["default_costs"])
    assert json.dumps(T.USER_COSTS) == json.dumps(golden["user_costs"])
    T.write_cost_files(str(tmp_path))
    assert json.load(open(os.path.join(str(tmp_path), "costs.json"))) == golden["default_costs"]
    assert json.load(open(os.path.join(str(tmp_path), "user_costs.json"))) == golden["user_costs"]


def test_distance_all_cases(golden):
    for c in all_cases(golden):
        d = O.distance(c["a"], c["b"], costs_of(golden, c["user"]))
        assert d == c["distance"], (c["a"], c["b"], d, c["distance"])


def test_distance_xml_600_pairs(golden):
    seqs = golden["xml_seqs"]
    for i, j, dd, du in golden["xml_all_pairs"]:
        assert O.distance(seqs[i], seqs[j], golden["default_costs"]) == dd
        assert O.distance(seqs[i], seqs[j], golden["user_costs"]) == du


def test_g6_fp64_fingerprints(golden):
    want = {100: 66.74999999999994, 200: 125.39999999999989, 300: 191.40999999999983,
            500: 315.9600000000003, 1000: 628.7499999999999}
    for c in golden["G6"]:
        assert c["distance"] == want[c["L"]]
        assert O.distance(c["a"], c["b"], golden["default_costs"]) == c["distance"]


def test_matrix_values_and_masks(golden):
    n = 0
    for c in all_cases(golden):
        if "dp_repr" not in c:
            continue
        D, M = O.matrix(c["a"], c["b"], costs_of(golden, c["user"]))
        want = np.array([[float(x) for x in row] for row in c["dp_repr"]])
        assert np.array_equal(D, want)
        assert np.array_equal(M, np.array(c["mask"], dtype=np.uint8))
        n += 1
    assert n > 200


def test_py_matrix_typing(golden):
    """int 0 vs 0.0 (SURVEY 8b 'value typing quirk') — str() parity with the reference."""
    for c in [golden["G1"]] + golden["small"][:80] + golden["empties"] :
        costs = costs_of(golden, c.get("user", False))
        D, M = O.py_matrix(c["a"], c["b"], costs)
        assert [[repr(x) for x in row] for row in D] == c["dp_repr"]
        assert M == c["mask"]
    assert str(O.py_matrix("AGRGA", "AGGGAA", golden["user_costs"])[0]) == golden["G1"]["dp_str"]


def test_all_paths_order_and_scripts(golden):
    n = 0
    for c in all_cases(golden):
        if "paths_cells" not in c:
            continue
        _, M = O.matrix(c["a"], c["b"], costs_of(golden, c["user"]))
        keep = len(c["paths_cells"])
        paths = O.all_paths(M.tolist(), cap=None if c["n_paths"] <= 4000 else keep)
        assert len(paths) == c["n_paths"]
        assert [[list(x) for x in p] for p in paths[:keep]] == c["paths_cells"]
        ess = [O.es_from_cells(p, c["a"], c["b"]) for p in paths[:keep]]
        assert ess[:3] == c["es"]
        assert [O.format_es(e) for e in ess] == c["es_fmt"]
        n += 1
    assert n > 250


def test_canonical_script_is_paths0(golden):
    for c in all_cases(golden):
        if "es" not in c:
            continue
        ops, oi, oj, d = O.canonical_script(c["a"], c["b"], costs_of(golden, c["user"]))
        assert d == c["distance"]
        assert O.es_from_ops(ops, oi, oj, c["a"], c["b"]) == c["es"][0]


def test_patch_rev_roundtrip(golden):
    for c in all_cases(golden):
        if "es" not in c:
            continue
        es0 = c["es"][0]
        assert list(O.patch_sequential(es0, c["a"])) == c["patch0"]
        assert list(O.patch_closed(es0, c["a"])) == c["patch0"]
        rev = O.rev_es(es0)
        assert rev == c["rev0"]
        assert list(O.patch_sequential(rev, c["b"])) == c["rev_patch0"]
        assert list(O.patch_closed(rev, c["b"])) == c["rev_patch0"]
        assert O.seq_from_es(es0) == c["seq_from_es0"]


def test_patch_error_codes_g1(golden):
    es1 = golden["G1"]["es"][0]
    want = {"AGRGA": [0, "AGGGAA"], "AGRGC": [1, "AGGGAA"], "AGRGAUU": [1, "AGGGAAUU"], "AGR": [-1, ""]}
    for x, out in golden["G1"]["patch_cases"]:
        assert list(O.patch_sequential(es1, x)) == out
        if x in want:
            assert out == want[x]
        assert list(O.patch_closed(es1, x)) == out


def test_patch_sequential_on_hand_edited_scripts(golden):
    for c in golden["patch_odd"]:
        assert list(O.patch_sequential(c["es"], c["x"])) == c["out"]


def test_patch_closed_codes_c(golden):
    for c in golden["small"][:120]:
        if "es" not in c:
            continue
        ops, oi, oj, _ = O.canonical_script(c["a"], c["b"], costs_of(golden, c["user"]))
        for x in (c["a"], c["a"] + "GU", c["a"][:-1] + ("G" if c["a"][-1] != "G" else "A"), c["a"][:-1]):
            want = O.patch_sequential(c["es"][0], x)
            code, out = O.patch_closed_codes(ops, oi, oj, O.encode(c["a"]), O.encode(c["b"]),
                                             O.encode(x) if x else np.zeros(0, np.uint8))
            assert (code, O.decode(out)) == want


def test_named_goldens_survey_appendix_b(golden):
    n = golden["xml_named"]
    assert n["3->2:default"]["distance"] == 1.0 and n["3->2:default"]["es_fmt"][0] == "[Upd(28,G)]"
    assert n["1->2:default"]["distance"] == 12.0 and n["1->2:default"]["n_paths"] == 56
    assert n["1->2:default"]["es_fmt"][0] == ("[Upd(12,U),Upd(14,C),Upd(15,U),Ins(15,A),Upd(19,C),Upd(20,A),"
                                              "Upd(22,U),Upd(23,U),Upd(24,G),Upd(27,G),Upd(28,G),Upd(29,U)]")
    assert n["1->2:user"]["distance"] == 16.0 and n["1->2:user"]["n_paths"] == 2
    assert n["12->13:default"]["es_fmt"][0] == "[Upd(29,U)]"
    assert n["24->25:user"]["distance"] == 12.0 and n["24->25:user"]["n_paths"] == 1


def test_search_topk_g7(golden):
    codes, off = O.concat(golden["xml_seqs"])
    for c in golden["G7"]:
        idx, sc, allsc = O.search_topk(c["query"], codes, off, golden["default_costs"], 6, want_scores=True)
        assert [s for _, s in c["scores"]] == allsc.tolist()
        assert [golden["xml_seqs"][i] for i in idx] == [s for s, _ in c["top6"]]
        assert sc.tolist() == [v for _, v in c["top6"]]
    g = golden["G7"][0]["top6"]
    assert [v for _, v in g] == [1.0, 0.1111111111111111, 0.1, 0.09090909090909091, 0.09090909090909091,
                                 0.08333333333333333]


def test_batch_matches_single(golden):
    cs = [c for c in golden["medium"]]
    a, ao = O.concat([c["a"] for c in cs]); b, bo = O.concat([c["b"] for c in cs])
    for user in (False, True):
        d = O.distance_batch(a, ao, b, bo, costs_of(golden, user), nthreads=4)
        for k, c in enumerate(cs):
            assert d[k] == O.distance(c["a"], c["b"], costs_of(golden, user))
