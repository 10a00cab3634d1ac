"""measures.py — host-side representations and single-pair measures of the reference's IRMethods.py.

These are the small objects the reference's GUI builds for ONE pair of sequences (gui.py:455-500: Python sets,
4-element multiset vectors, 15x15 bigram matrices) and the twelve measures it applies to them (IR:49-389).
They are not the hot path (the collection scan of the same measures runs on the GPU,
k_sim.cuh); they exist so that dropin/IRMethods.py satisfies every importer of the reference module
(gui.py:18-22, timing.py:4, fa_import.py:5) with identical values.  Everything is written from the
reference's observable behaviour: per-cell accumulation in sequence order, numpy reductions where the
reference uses numpy reductions (np.sum is pairwise — the same calls give the same bits).

Table-driven: one 15x4 weight matrix replaces the reference's two ambiguity dictionaries (IR:20-46)."""
from __future__ import annotations

import math
import pickle
import time

import numpy as np

nucleotides = ['A', 'G', 'C', 'U', 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N']   # IR:13
base_nucleotides = nucleotides[:4]                                                          # IR:15

_AMBIG = {"Y": (0, 0, .5, .5), "R": (.5, .5, 0, 0), "W": (.5, 0, 0, .5), "S": (0, .5, .5, 0), "K": (0, .5, 0, .5),
          "M": (.5, 0, .5, 0), "D": (.33, .33, 0, .33), "V": (.33, .33, .33, 0), "H": (.33, 0, .33, .33),
          "B": (0, .33, .33, .33), "N": (.25, .25, .25, .25)}
# the two dictionaries importers of the reference module can see (IR:20-46)
ambiguous_nucleotides = {s: np.array(w, dtype=float) for s, w in _AMBIG.items()}
ambiguity_vectors = {s: dict(zip(base_nucleotides, w)) for s, w in _AMBIG.items()}

_CODE = {s: i for i, s in enumerate(nucleotides)}
_W = np.zeros((15, 4))
_W[:4] = np.eye(4)
for _s, _w in _AMBIG.items():
    _W[_CODE[_s]] = _w


def _codes(seq):
    try:
        return [_CODE[c] for c in seq]
    except KeyError as e:            # the reference: ValueError from list.index (IR:155) / KeyError (IR:106)
        raise ValueError(f"{e.args[0]!r} is not in list") from None


# ---- representations -----------------------------------------------------------------------------------
def convert_to_set(sequence):
    """IR:49-51."""
    return set(sequence)


def convert_to_multi_set(sequence):
    """IR:95-107: expected base counts (A, G, C, U), symbols added one at a time in sequence order."""
    c = np.zeros(4)
    for ch in sequence:
        if ch in base_nucleotides:
            c[_CODE[ch]] += 1
        else:
            c = c + ambiguous_nucleotides[ch]            # KeyError for unknown symbols, like IR:106
    return c


def convert_to_tf_vector(seq):
    """IR:147-186: 15x15 bigram counts; a bigram holding an ambiguity code also credits the base bigrams it may
    stand for (probability products).  Every cell receives its addends in sequence order."""
    vec = np.zeros((15, 15))
    cs = _codes(seq)
    for cur, nxt in zip(cs[:-1], cs[1:]):
        vec[cur][nxt] += 1
        if nxt >= 4:
            vec[:4, :4] += np.outer(_W[cur], _W[nxt])
        elif cur >= 4:
            vec[:4, nxt] += _W[cur]
    return vec


def get_base_possibilities(base):
    """IR:277-287: the base itself (probability 1) and every ambiguity code that can stand for it."""
    syms, probs = [base], [1]
    for s in ambiguity_vectors:
        if ambiguity_vectors[s][base] != 0:
            syms.append(s)
            probs.append(ambiguity_vectors[s][base])
    return syms, probs


def possibilities(nucleotide):
    """IR:255-274."""
    if nucleotide in base_nucleotides:
        return get_base_possibilities(nucleotide)
    syms = [b for b in base_nucleotides if ambiguity_vectors[nucleotide][b] != 0]
    probs = [1 / len(syms)] * len(syms)
    return syms + [nucleotide], probs + [1]


def compare_pair_to_seq(pair, record, is_document=True):
    """IR:225-252: 1 when the bigram occurs literally, else the best probability product over the bigrams it may
    stand for that do occur."""
    seq = record['sequence'] if is_document else record
    if pair in seq:
        return 1
    p1, w1 = possibilities(pair[0])
    p2, w2 = possibilities(pair[1])
    best = 0
    for a, wa in zip(p1, w1):
        for b, wb in zip(p2, w2):
            if a + b in seq and best < wa * wb:
                best = wa * wb
    return best


def convert_to_idf_vector(seq1, collection=None, list_of_docs=None, doc_count=0):
    """IR:189-216: log10(N / weighted document frequency) for every distinct bigram of seq1."""
    vec = np.zeros((15, 15))
    pairs = {seq1[i] + seq1[i + 1] for i in range(len(seq1) - 1)}
    count = doc_count if doc_count != 0 else len(list_of_docs) if list_of_docs is not None else collection.count_documents({})
    for pair in pairs:
        df = 0
        for record in (list_of_docs if list_of_docs is not None else collection.find({})):
            df = df + compare_pair_to_seq(pair, record, is_document=collection is not None)
        vec[nucleotides.index(pair[0])][nucleotides.index(pair[1])] = 0 if df == 0 else math.log(count / df, 10)
    return vec


def create_tf_idf_vector(seq, collection=None, list_of_docs=None, doc_count=0, is_document=False):
    """IR:219-223: the matrix product tf . idf (np.dot, as the reference)."""
    if is_document:
        return np.dot(pickle.loads(seq['tf']), pickle.loads(seq['idf']))
    return np.dot(convert_to_tf_vector(seq), convert_to_idf_vector(seq, collection, list_of_docs, doc_count))


# ---- measures on pre-built representations ----------------------------------------------------------
def _deliver(key):
    """The reference's calling convention: return the value, or store it under `key` when a dict is passed."""
    def deco(fn):
        def wrapper(a, b, return_dict=None):
            val = fn(a, b)
            if return_dict is None:
                return val
            return_dict[key] = val
        wrapper.__name__ = wrapper.__qualname__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        wrapper.return_key = key
        return wrapper
    return deco


@_deliver('intersection')
def intersection(a, b):
    """IR:54-60."""
    return a.intersection(b)


@_deliver('set_intersection_sim')
def set_intersection_similarity(a, b):
    """IR:63-69."""
    return len(a.intersection(b))


@_deliver('set_jaccard_sim')
def set_jaccard_similarity(a, b):
    """IR:72-79."""
    return len(a.intersection(b)) / len(a.union(b))


@_deliver('set_dice_sim')
def set_dice_similarity(a, b):
    """IR:82-91."""
    return 2 * len(a.intersection(b)) / (len(a) + len(b))


def _multi_min(ca, cb):
    sim = 0
    for k in range(4):
        sim += min(ca[k], cb[k])
    return sim


@_deliver('multi_intersection_sim')
def multi_intersection_similarity(ca, cb):
    """IR:110-121."""
    return _multi_min(ca, cb)


@_deliver('multi_jaccard_sim')
def multi_jaccard_similarity(ca, cb):
    """IR:124-132."""
    num = _multi_min(ca, cb)
    return num / (np.sum(ca) + np.sum(cb) - num)


@_deliver('multi_dice_sim')
def multi_dice_similarity(ca, cb):
    """IR:135-145."""
    num = _multi_min(ca, cb)
    return 2 * num / (np.sum(ca) + np.sum(cb))


def _dot_and_norms(a, b):
    return np.sum(np.multiply(a, b)), np.sum(np.square(a)), np.sum(np.square(b))


@_deliver('cosine')
def cosine(a, b):
    """IR:290-304."""
    num, a_sq, b_sq = _dot_and_norms(a, b)
    return num / math.sqrt(a_sq * b_sq)


@_deliver('pearson')
def pearson(a, b):
    """IR:307-329."""
    num, a_sq, b_sq = _dot_and_norms(np.subtract(a, np.average(a)), np.subtract(b, np.average(b)))
    return num / math.sqrt(a_sq * b_sq)


@_deliver('euclidian_dist')
def euclidian_distance(a, b):
    """IR:332-340."""
    return 1 / (1 + math.sqrt(np.sum(np.square(np.subtract(a, b)))))


@_deliver('manhattan_distance')
def manhattan_distance(a, b):
    """IR:343-351 (the reference takes a square root here too)."""
    return 1 / (1 + math.sqrt(np.sum(np.abs(np.subtract(a, b)))))


@_deliver('tanimoto_dist')
def tanimoto_distance(a, b):
    """IR:354-369."""
    num, a_sq, b_sq = _dot_and_norms(a, b)
    return num / (a_sq + b_sq - num)


@_deliver('dice_dist')
def dice_dist(a, b):
    """IR:372-389."""
    num, a_sq, b_sq = _dot_and_norms(a, b)
    return 2 * num / (a_sq + b_sq)


SET_MEASURES = (set_intersection_similarity, set_jaccard_similarity, set_dice_similarity)
MULTI_MEASURES = (multi_intersection_similarity, multi_jaccard_similarity, multi_dice_similarity)
VECTOR_MEASURES = (cosine, pearson, euclidian_distance, manhattan_distance, tanimoto_distance, dice_dist)


def time_method(method):
    """IR:392-399: run `method` and also store its wall time (ms) under '<name>_time'."""
    def wrapper(a, b, return_dict):
        start = time.time()
        method(a, b, return_dict)
        return_dict[method.__name__ + '_time'] = (time.time() - start) * 1000
    return wrapper


def create_and_start_threads(methods_to_execute, a, b):
    """IR:402-419.  The reference forks one process per method and collects into a Manager dict; the measures
    are microseconds of numpy on one pair, so they simply run in turn here (no fork next to a CUDA context) and
    the same keys come back in a plain dict."""
    return_dict = {}
    for m in methods_to_execute:
        time_method(m)(a, b, return_dict)
    return return_dict


def perform_methods(a, b, do_cosine=False, do_pearson=False, do_euclidian_distance=False, do_manhattan_distance=False,
                    do_tanimoto_distance=False, do_dice_dist=False):
    """IR:421-432."""
    wanted = (do_cosine, do_pearson, do_euclidian_distance, do_manhattan_distance, do_tanimoto_distance, do_dice_dist)
    return create_and_start_threads([m for m, on in zip(VECTOR_MEASURES, wanted) if on], a, b)
