"""ir.py — the search side of the reference's IRMethods.py on top of the CUDA engine.

Mirrors: wf_score IR:435-440, the wf_score branch of search_collection IR:443-447,466-477 (one
batched GPU scan of the collection instead of one wagnerFisher call per document), and the
stable descending top-k of performance.py:12-15 / gui.py:573,593; and the other scorers of
search_collection — set / multiset / TF-vector measures, IR:49-389 — as one GPU pass over the collection
(similarity_collection)."""
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


SIM_METHODS = ("set_intersection_similarity", "set_jaccard_similarity", "set_dice_similarity",
               "multi_intersection_similarity", "multi_jaccard_similarity", "multi_dice_similarity",
               "cosine", "pearson", "euclidian_distance", "manhattan_distance", "tanimoto_distance", "dice_dist")


def similarity_collection(query: str, sequences, method: str, engine=None):
    """[(sequence, method(query, sequence))] in collection order for one of SIM_METHODS — what
    search_collection(query, 'tf', collection, <method>) returns (IR:449-477), one GPU pass.  The documents'
    stored 'tf' vectors are rebuilt from the sequences on the device (they are convert_to_tf_vector(sequence),
    fa_import.py:49).  Unknown symbols raise KeyError (the reference: KeyError IR:106 / ValueError IR:155)."""
    if method not in SIM_METHODS:
        raise ValueError(f"unknown similarity method {method!r}")
    sequences = list(sequences)
    if not sequences:
        return []
    from .encoding import encode
    eng = engine or get_engine()
    qc = encode(query)
    eng.db_load(pack(sequences, bits=4))
    try:
        scores, _, _ = eng.db_similarity(qc, method)
    finally:
        eng.db_free()
    return [(s, float(v)) for s, v in zip(sequences, scores)]


def top_k(scores, k):
    """performance.py:12-15 — stable, so ties keep collection order."""
    return sorted(scores, key=itemgetter(1), reverse=True)[0:k]
