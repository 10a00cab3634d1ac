"""The drop-in modules satisfy every importer of the reference (CPU: no compute call touches the GPU here).

Runs the exact import lists of /root/reference/gui.py:13,18-22, timing.py:1,4, performance.py:3 and fa_import.py:5
against rna-sequence-diff-patch_b200/dropin/, and checks the host-side representations and measures
(rna_sequence_diff_patch_b200/measures.py) bit for bit against what the unmodified reference produced
(tests/golden/ref_golden_ir.json, made by tests/golden/make_golden_ir.py)."""
import json
import math
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROP = os.path.join(ROOT, "rna-sequence-diff-patch_b200", "dropin")


@pytest.fixture(scope="module")
def gir():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_golden_ir.json")))


@pytest.fixture(scope="module")
def IR():
    sys.path.insert(0, DROP)
    try:
        import IRMethods
        return IRMethods
    finally:
        sys.path.remove(DROP)


def unhex(xs):
    return np.array([math.nan if x == "nan" else float.fromhex(x) for x in xs], dtype=np.float64)


def same(a, b):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    return (np.array_equal(np.isnan(a), np.isnan(b))
            and np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)]))


def _run(code, cwd):
    env = dict(os.environ, PYTHONPATH=DROP)
    return subprocess.run([sys.executable, "-c", textwrap.dedent(code)], cwd=cwd, env=env, capture_output=True, text=True)


def test_importers_of_the_reference_find_every_name(tmp_path):
    """The import statements of the reference's own callers, verbatim, in a directory without cost files."""
    r = _run("""
        from StringEditDistance import wagnerFisher, create_paths, generate_es, patching, generate_rev_es, reload_user_costs, user_costs
        from IRMethods import (cosine, pearson, euclidian_distance, manhattan_distance, tanimoto_distance, dice_dist,
                               set_intersection_similarity, set_dice_similarity, set_jaccard_similarity,
                               multi_intersection_similarity, multi_dice_similarity, multi_jaccard_similarity, convert_to_set,
                               convert_to_tf_vector, convert_to_idf_vector, create_tf_idf_vector,
                               convert_to_multi_set, create_and_start_threads, wf_score, create_search_threads, search_collection)
        from IRMethods import search_collection, wf_score, cosine, pearson          # performance.py:3
        from IRMethods import convert_to_tf_vector, convert_to_idf_vector           # fa_import.py:5
        import StringEditDistance as S
        assert S.user_costs is S.default_costs                                      # SED:16
        ns = {}
        exec("from StringEditDistance import wagnerFisher\\nfrom IRMethods import *", ns)   # timing.py:1,4
        for name in ("convert_to_set", "convert_to_multi_set", "convert_to_tf_vector", "convert_to_idf_vector",
                     "set_intersection_similarity", "set_jaccard_similarity", "set_dice_similarity",
                     "multi_intersection_similarity", "multi_jaccard_similarity", "multi_dice_similarity",
                     "cosine", "pearson", "euclidian_distance", "manhattan_distance", "tanimoto_distance", "dice_dist",
                     "nucleotides", "base_nucleotides", "ambiguous_nucleotides", "ambiguity_vectors", "perform_methods",
                     "time_method", "compare_pair_to_seq", "possibilities", "get_base_possibilities", "intersection",
                     "np", "math", "pickle", "time", "itemgetter", "os", "Process", "Manager"):
            assert name in ns, name
        print("ok")
    """, str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert r.stdout.split() == ["Could", "not", "find", "user", "costs", "file", "ok"]


def test_user_costs_file_is_read_from_cwd(tmp_path, golden):
    (tmp_path / "user_costs.json").write_text(json.dumps(golden["user_costs"]))
    (tmp_path / "costs.json").write_text(json.dumps(golden["default_costs"]))
    r = _run("""
        import StringEditDistance as S
        assert S.user_costs is not S.default_costs and S.user_costs["insert"] == 2.0 and S.default_costs["insert"] != 2.0
        S.user_costs["insert"] = 5
        import json; json.dump(S.user_costs, open("user_costs.json", "w"))
        S.reload_user_costs()
        assert S.user_costs["insert"] == 5
        print("ok")
    """, str(tmp_path))
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_representations_match_reference(IR, gir):
    for s, v in gir["tf"].items():
        assert same(IR.convert_to_tf_vector(s), unhex(v)), s
    for s, v in gir["multiset"].items():
        assert same(IR.convert_to_multi_set(s), unhex(v)), s
    assert IR.convert_to_set("AGGN") == {"A", "G", "N"}
    for c, (syms, probs) in gir["possibilities"].items():
        got = IR.possibilities(c)
        assert got[0] == syms and same(np.array(got[1], float), unhex(probs)), c
    for p, s, v in gir["compare_pair"]:
        assert same(np.array([IR.compare_pair_to_seq(p, s, is_document=False)], float), unhex([v])), (p, s)
    with pytest.raises(ValueError):
        IR.convert_to_tf_vector("AXG")
    with pytest.raises(KeyError):
        IR.convert_to_multi_set("AXG")


def test_measures_on_prebuilt_representations_match_reference(IR, gir):
    """What gui.py:462-500 does: build the sets / vectors, then call the measure objects on them."""
    docs = gir["docs"]
    for q in (gir["queries"][1], gir["queries"][4]):
        for name in gir["methods"]:
            conv = IR.convert_to_set if name.startswith("set_") else IR.convert_to_multi_set if name.startswith("multi_") else IR.convert_to_tf_vector
            a = conv(q)
            want = unhex(gir["scores"][q][name])
            with np.errstate(all="ignore"):
                got = []
                for d in docs:
                    try:
                        got.append(float(getattr(IR, name)(a, conv(d))))
                    except ZeroDivisionError:             # int / int in the set measures; numpy 0/0 elsewhere is nan
                        got.append(math.nan)
            assert same(np.array(got), want), (q, name)
    rd = {}
    IR.cosine(IR.convert_to_tf_vector("ACGU"), IR.convert_to_tf_vector("ACGA"), rd)
    assert list(rd) == ["cosine"]
    out = IR.create_and_start_threads([IR.cosine, IR.dice_dist], IR.convert_to_tf_vector("ACGU"), IR.convert_to_tf_vector("ACGA"))
    assert set(out) == {"cosine", "dice_dist", "cosine_time", "dice_dist_time"} and out["cosine"] == rd["cosine"]
    out = IR.perform_methods(IR.convert_to_tf_vector("ACGU"), IR.convert_to_tf_vector("ACGA"), do_pearson=True)
    assert set(out) == {"pearson", "pearson_time"}


class _Coll:
    def __init__(self, docs):
        self.docs = docs

    def find(self, _):
        return iter(self.docs)

    def count_documents(self, _):
        return len(self.docs)


def test_idf_and_tfidf_vectors_and_search_match_reference(IR, gir):
    small = gir["idf_docs"]
    coll = _Coll([{"sequence": s} for s in small])
    for s, v in gir["idf"].items():
        assert same(IR.convert_to_idf_vector(s, coll), unhex(v)), s
    for key, rec in gir["idf_pairlist"].items():
        a, b = key.split("|")
        assert same(IR.convert_to_idf_vector(a, list_of_docs=[a, b]), unhex(rec["idf_a"]))
        assert same(IR.create_tf_idf_vector(a, list_of_docs=[a, b]), unhex(rec["tfidf_a"]))
        assert same(IR.create_tf_idf_vector(b, list_of_docs=[a, b]), unhex(rec["tfidf_b"]))
    # documents with the stored vectors (what the reference needs) and documents without (computed from the sequence)
    stored = _Coll([{"sequence": s, "tf": pickle.dumps(IR.convert_to_tf_vector(s)),
                     "idf": pickle.dumps(IR.convert_to_idf_vector(s, coll))} for s in small])
    with np.errstate(all="ignore"):
        for q, per_vt in gir["idf_scores"].items():
            for vt, per_m in per_vt.items():
                for name, want in per_m.items():
                    for c in (stored, coll):
                        res = IR.search_collection(q, vt, c, getattr(IR, name))
                        assert [s for s, _ in res] == small
                        assert same(np.array([v for _, v in res], float), unhex(want)), (q, vt, name)
