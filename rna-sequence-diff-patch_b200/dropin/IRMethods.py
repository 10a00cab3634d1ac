"""Drop-in for the search side of the reference's IRMethods.py: wf_score (IR:435-440) and
search_collection (IR:443-477) with method == wf_score or one of the set / multiset / TF-vector measures
(IR:49-389, vector_type 'tf' — the only vector the reference's importer stores, fa_import.py:49).  Every
method is one GPU pass over the collection.  The measure functions below are the objects callers pass as
`method`; called directly on two SEQUENCES they score that pair on the GPU (the reference calls them on
pre-built sets / vectors, which this module does not build on the host)."""
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


def _measure(name):
    def fn(a, b, return_dict=None):
        if not (isinstance(a, str) and isinstance(b, str)):
            raise TypeError(name + ": pass two sequences (the GPU builds the sets / vectors itself)")
        val = _ir.similarity_collection(a, [b], name)[0][1]
        if return_dict is None:
            return val
        return_dict[_RETURN_KEYS[name]] = val
    fn.__name__ = fn.__qualname__ = name
    return fn


# keys under which the reference's measures store their result in return_dict (IR:69,79,91,121,132,145,304,...)
_RETURN_KEYS = {"set_intersection_similarity": "set_intersection_sim", "set_jaccard_similarity": "set_jaccard_sim",
                "set_dice_similarity": "set_dice_sim", "multi_intersection_similarity": "multi_intersection_sim",
                "multi_jaccard_similarity": "multi_jaccard_sim", "multi_dice_similarity": "multi_dice_sim",
                "cosine": "cosine", "pearson": "pearson", "euclidian_distance": "euclidian_dist",
                "manhattan_distance": "manhattan_distance", "tanimoto_distance": "tanimoto_dist", "dice_dist": "dice_dist"}
set_intersection_similarity = _measure("set_intersection_similarity")
set_jaccard_similarity = _measure("set_jaccard_similarity")
set_dice_similarity = _measure("set_dice_similarity")
multi_intersection_similarity = _measure("multi_intersection_similarity")
multi_jaccard_similarity = _measure("multi_jaccard_similarity")
multi_dice_similarity = _measure("multi_dice_similarity")
cosine = _measure("cosine")
pearson = _measure("pearson")
euclidian_distance = _measure("euclidian_distance")
manhattan_distance = _measure("manhattan_distance")
tanimoto_distance = _measure("tanimoto_distance")
dice_dist = _measure("dice_dist")
_SIM = {f: f.__name__ for f in (set_intersection_similarity, set_jaccard_similarity, set_dice_similarity,
                                multi_intersection_similarity, multi_jaccard_similarity, multi_dice_similarity,
                                cosine, pearson, euclidian_distance, manhattan_distance, tanimoto_distance, dice_dist)}


def search_collection(query, vector_type, collection, method, return_dict=None, callback=None):
    """wf_score (always with the default costs, IR:470) or one of the measures above; vector measures use the
    'tf' vectors (IR:455-457) — 'idf' / 'tf-idf' documents do not exist in the reference's database either."""
    docs = [doc['sequence'] for doc in collection.find({})]
    if method is wf_score:
        scores = _ir.score_collection(query, docs, _SED.default_costs)
    elif method in _SIM:
        if _SIM[method] in _ir.SIM_METHODS[6:] and vector_type != 'tf':
            raise NotImplementedError("vector measures are served for vector_type 'tf' only")
        scores = _ir.similarity_collection(query, docs, _SIM[method])
    else:
        raise NotImplementedError("search_collection: unknown method " + getattr(method, '__name__', repr(method)))
    if callback is not None:
        callback(scores)
    elif return_dict is not None:
        return_dict[method.__name__] = scores
    else:
        return scores


def create_search_threads(methods_to_execute, query, vector_type, collection, on_search_done=None, on_wf_done=None):
    """IR:480-515: one GPU scan of the collection per method instead of a forked process per method plus
    a second, separate WF pass (IR:487-491,511-515); no pandas (the reference's DataFrame.append, IR:501, no
    longer exists).  The callbacks receive what the reference passes them: a list of (sequence, mean score
    over the methods) — duplicates collapse like the reference's dict, the mean skips NaN like
    DataFrame.mean (IR:501-505) — and the raw wf_score list."""
    per_method = [search_collection(query, vector_type, collection, m) for m in methods_to_execute]
    if on_search_done is not None:
        order, rows = [], []
        for scores in per_method:
            row = {}
            for seq, sc in scores:
                if seq not in row and seq not in order:
                    order.append(seq)
                row[seq] = sc
            rows.append(row)
        final = []
        for seq in order:
            vals = [r[seq] for r in rows if seq in r and r[seq] == r[seq]]
            final.append((seq, sum(vals) / len(vals) if vals else float('nan')))
        on_search_done(final)
    if wf_score in methods_to_execute and on_wf_done is not None:
        on_wf_done(per_method[list(methods_to_execute).index(wf_score)])


def top_k(scores, k):
    return _ir.top_k(scores, k)
