"""Drop-in for the Wagner–Fischer part of the reference's IRMethods.py: wf_score (IR:435-440) and
search_collection with method == wf_score (IR:443-447,466-477).  The other similarity measures of
the reference (IR:49-389) are outside this package's scope."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import StringEditDistance as _SED  # noqa: E402  (the drop-in next to this file)
from StringEditDistance import wagnerFisher  # noqa: E402,F401
from rna_sequence_diff_patch_b200 import ir as _ir  # noqa: E402

nucleotides = ['A', 'G', 'C', 'U', 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N']


def wf_score(seq1, seq2, user_cost=False):
    return _ir.wf_score(seq1, seq2, _SED.user_costs if user_cost else _SED.default_costs)


def search_collection(query, vector_type, collection, method, return_dict=None, callback=None):
    """Only method == wf_score is served (always with the default costs, IR:470)."""
    if method is not wf_score:
        raise NotImplementedError("this drop-in serves search_collection(..., wf_score) only")
    docs = [doc['sequence'] for doc in collection.find({})]
    scores = _ir.score_collection(query, docs, _SED.default_costs)
    if callback is not None:
        callback(scores)
    elif return_dict is not None:
        return_dict[method.__name__] = scores
    else:
        return scores


def create_search_threads(methods_to_execute, query, vector_type, collection, on_search_done=None, on_wf_done=None):
    """IR:480-515 for the Wagner-Fischer method: ONE scan of the collection on the GPU instead of the
    reference's forked process per method plus a second, separate WF pass (IR:487-491,511-515); no
    pandas (the reference's DataFrame.append, IR:501, no longer exists).  The callbacks receive what
    the reference passes them: a list of (sequence, mean score) and the raw wf_score list."""
    others = [m for m in methods_to_execute if m is not wf_score]
    if others:
        raise NotImplementedError("this drop-in serves the wf_score search only")
    scores = search_collection(query, vector_type, collection, wf_score)
    if on_search_done is not None:
        merged = {}
        for seq, sc in scores:                      # duplicates collapse like the reference's dict (IR:501)
            merged[seq] = sc
        on_search_done(list(merged.items()))
    if wf_score in methods_to_execute and on_wf_done is not None:
        on_wf_done(scores)


def top_k(scores, k):
    return _ir.top_k(scores, k)
