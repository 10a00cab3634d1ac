"""ir.py — the Wagner–Fischer part of the reference's IRMethods.py on top of the CUDA engine.

Mirrors: wf_score IR:435-440, the wf_score branch of search_collection IR:443-447,466-477 (one
batched GPU scan of the collection instead of one wagnerFisher call per document), and the
stable descending top-k of performance.py:12-15 / gui.py:573,593."""
from __future__ import annotations

from operator import itemgetter

import numpy as np

from . import sed
from .encoding import pack
from .engine import get_engine


def wf_score(seq1: str, seq2: str, costs: dict) -> float:
    """IR:435-440: 1 / (1 + D[m][n])."""
    return 1 / (1 + sed.distance(seq1, seq2, costs))


def score_collection(query: str, sequences, costs: dict, engine=None):
    """[(sequence, wf_score(query, sequence))] in collection order — what
    search_collection(query, _, collection, wf_score) returns (IR:469-477), one GPU pass."""
    sequences = list(sequences)
    if not sequences:
        return []
    eng = engine or get_engine()
    eng.set_costs(costs)
    for s in sequences:                       # the reference's per-document KeyError (SED:87)
        sed._validate_and_encode(query, s, costs)
    up = [s.upper() for s in sequences]
    eng.db_load(pack(up, bits=4))
    try:
        _, _, scores = eng.db_search_topk(pack([query.upper()], bits=4), k=1, want_scores=True)
    finally:
        eng.db_free()
    return [(s, float(v)) for s, v in zip(sequences, scores[0])]


def top_k(scores, k):
    """performance.py:12-15 — stable, so ties keep collection order."""
    return sorted(scores, key=itemgetter(1), reverse=True)[0:k]
