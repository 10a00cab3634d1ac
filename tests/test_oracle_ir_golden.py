"""Pins oracle/ir_oracle.py (set / multiset / TF-vector similarities) to the outputs of
the unmodified reference recorded in tests/golden/ref_golden_ir.json — bit for bit."""
import json
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ir_oracle as IO  # noqa: E402


@pytest.fixture(scope="module")
def gir():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "ref_golden_ir.json")))


def unhex(xs):
    return np.array([math.nan if x == "nan" else float.fromhex(x) for x in xs], dtype=np.float64)


def same(a, b):
    return np.array_equal(a.view(np.uint64)[~np.isnan(a)], b.view(np.uint64)[~np.isnan(b)]) and np.array_equal(np.isnan(a), np.isnan(b))


def test_representations_match_reference(gir):
    for s, v in gir["tf"].items():
        assert same(IO.tf_vector(s).reshape(-1), unhex(v)), s
    for s, v in gir["multiset"].items():
        assert same(IO.multiset(s), unhex(v)), s


@pytest.mark.parametrize("method", IO.METHODS)
def test_scores_match_reference(gir, method):
    assert gir["methods"] == IO.METHODS
    for q in gir["queries"]:
        got = IO.search(q, gir["docs"], method)
        assert same(got, unhex(gir["scores"][q][method])), (q, method)
