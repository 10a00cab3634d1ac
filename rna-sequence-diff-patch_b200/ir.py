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
from .engine import get_engine, get_search_engine


def wf_score(seq1: str, seq2: str, costs: dict) -> float:
    """IR:435-440: 1 / (1 + D[m][n])."""
    return 1 / (1 + sed.distance(seq1, seq2, costs))


def _validate_collection(query: str, sequences, costs: dict):
    """The reference raises KeyError for the first (document, row symbol, column symbol) whose update cost is not in
    the table (SED:87 through IR:437).  One pass over the distinct symbols of the whole collection decides whether
    any document can fail; only then is the per-document check (which finds the reference's first error) run."""
    upd = costs["update"]
    doc_syms = set().union(*map(set, sequences)) if len(sequences) < 64 else set("".join(sequences))
    clean = True
    for c1 in set(query):
        row = upd.get(c1)
        for c2 in doc_syms:
            if c2.lower() != c1.lower() and (row is None or c2 not in row):
                clean = False
    if not clean:
        for s in sequences:
            sed._validate_and_encode(query, s, costs)


class _Resident:
    """The collection last loaded by score_collection / similarity_collection: a session that searches the same
    collection again (gui.py:532-600: every search re-reads collection.find({})) pays rsd_db_load once."""
    engine = None
    sequences = None
    gen = -1

    @classmethod
    def load(cls, eng, sequences, upper: bool):
        if (cls.engine is eng and cls.gen == getattr(eng, "_db_gen", 0) and cls.sequences is not None
                and cls.sequences[0] == upper and cls.sequences[1] == sequences):
            return
        if cls.engine is not None and cls.engine is not eng:
            cls.drop()                                   # the collection moves to another engine: free the old copy
        cls.engine, cls.sequences = None, None
        eng.db_load(pack([s.upper() for s in sequences] if upper else sequences, bits=4))
        cls.engine, cls.sequences, cls.gen = eng, (upper, sequences), eng._db_gen

    @classmethod
    def drop(cls):
        if cls.engine is not None and cls.gen == getattr(cls.engine, "_db_gen", 0):
            cls.engine.db_free()
        cls.engine, cls.sequences = None, None


def release_collection():
    """Free the device copy of the collection kept by the last search (optional; the next search reloads)."""
    _Resident.drop()


def score_collection(query: str, sequences, costs: dict, engine=None):
    """[(sequence, wf_score(query, sequence))] in collection order — what
    search_collection(query, _, collection, wf_score) returns (IR:469-477): one pass over the collection, sharded
    over every visible GPU (one process, NCCL gather of the top-k lists; the per-document scores come back per shard)."""
    sequences = list(sequences)
    if not sequences:
        return []
    eng = engine or get_search_engine()
    eng.set_costs(costs)
    _validate_collection(query, sequences, costs)
    _Resident.load(eng, sequences, upper=True)
    _, _, scores = eng.db_search_topk(pack([query.upper()], bits=4), k=1, want_scores=True)
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
    _Resident.load(eng, sequences, upper=False)
    scores, _, _ = eng.db_similarity(qc, method)
    return [(s, float(v)) for s, v in zip(sequences, scores)]


def top_k(scores, k):
    """performance.py:12-15 — stable, so ties keep collection order."""
    return sorted(scores, key=itemgetter(1), reverse=True)[0:k]
