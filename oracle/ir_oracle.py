"""ir_oracle.py — CPU restatement (numpy) of the reference's set / multiset / TF-vector similarity
measures, SURVEY section 8(f) rank 4.  TEST INFRASTRUCTURE ONLY: imported by tests/ and nowhere in the
product path.  PINNED: tests/test_oracle_ir_golden.py checks every function here bit for bit against
tests/golden/ref_golden_ir.json, which tests/golden/make_golden_ir.py produced by running the
unmodified reference (IRMethods.search_collection) in the build container.

The reference computes these with numpy itself (np.sum = pairwise summation, np.average, np.multiply
...), so this restatement calls the same numpy reductions; the CUDA kernels re-create numpy's pairwise
summation order explicitly (see k_sim.cuh) and are compared with this file."""
from __future__ import annotations

import math

import numpy as np

SYMBOLS = "AGCUYRWSKMDVHBN"                      # IR:13 — index = code
# IR:20-32 / IR:34-46: probability of each base (A, G, C, U) behind every symbol; bases are unit vectors
W = np.zeros((15, 4))
W[:4] = np.eye(4)
for _s, _w in {"Y": (0, 0, .5, .5), "R": (.5, .5, 0, 0), "W": (.5, 0, 0, .5), "S": (0, .5, .5, 0), "K": (0, .5, 0, .5),
               "M": (.5, 0, .5, 0), "D": (.33, .33, 0, .33), "V": (.33, .33, .33, 0), "H": (.33, 0, .33, .33),
               "B": (0, .33, .33, .33), "N": (.25, .25, .25, .25)}.items():
    W[SYMBOLS.index(_s)] = _w

METHODS = ["set_intersection_similarity", "set_jaccard_similarity", "set_dice_similarity",
           "multi_intersection_similarity", "multi_jaccard_similarity", "multi_dice_similarity",
           "cosine", "pearson", "euclidian_distance", "manhattan_distance", "tanimoto_distance", "dice_dist"]


def codes_of(seq: str) -> list:
    return [SYMBOLS.index(c) for c in seq]


def set_mask(seq: str) -> int:
    """IR:49-51: set(sequence) as a 15-bit mask."""
    m = 0
    for c in codes_of(seq):
        m |= 1 << c
    return m


def multiset(seq: str) -> np.ndarray:
    """IR:95-107: per symbol, in sequence order, add the symbol's base weights to a 4-vector (fp64)."""
    c = np.zeros(4)
    for s in codes_of(seq):
        if s < 4:
            c[s] += 1
        else:
            c = c + W[s]
    return c


def tf_vector(seq: str) -> np.ndarray:
    """IR:147-186: 15x15 bigram matrix; ambiguity codes also credit the base bigrams they may stand for."""
    v = np.zeros((15, 15))
    cs = codes_of(seq)
    for cur, nxt in zip(cs[:-1], cs[1:]):
        v[cur][nxt] += 1                                      # IR:156
        wc = W[cur]
        if nxt < 4 and cur >= 4:                              # IR:171-174
            for k in range(4):
                v[k][nxt] += wc[k]
        elif nxt >= 4:                                        # IR:176-184
            for k in range(4):
                for j in range(4):
                    v[k][j] += wc[k] * W[nxt][j]
    return v


def represent(seq: str, method: str):
    if method.startswith("set_"):
        return set_mask(seq)
    if method.startswith("multi_"):
        return multiset(seq)
    return tf_vector(seq)


def score(method: str, a, b) -> float:
    """a = query representation, b = document representation (IR:469: method(vector1, convert(doc)))."""
    with np.errstate(all="ignore"):
        if method.startswith("set_"):
            inter = bin(a & b).count("1")                                        # IR:64-69
            if method == "set_intersection_similarity":
                return float(inter)
            if method == "set_jaccard_similarity":                                # IR:72-79
                return inter / bin(a | b).count("1")
            return 2 * inter / (bin(a).count("1") + bin(b).count("1"))           # IR:82-91
        if method.startswith("multi_"):
            sim = 0
            for k in range(4):                                                    # IR:110-116
                sim += min(a[k], b[k])
            if method == "multi_intersection_similarity":
                return float(sim)
            if method == "multi_jaccard_similarity":                              # IR:124-132
                return float(sim / (np.sum(a) + np.sum(b) - sim))
            return float(2 * sim / (np.sum(a) + np.sum(b)))                       # IR:135-145
        if method == "cosine":                                                    # IR:290-304
            return float(np.sum(np.multiply(a, b)) / math.sqrt(np.sum(np.square(a)) * np.sum(np.square(b))))
        if method == "pearson":                                                   # IR:307-329
            xa, xb = np.subtract(a, np.average(a)), np.subtract(b, np.average(b))
            return float(np.sum(np.multiply(xa, xb)) / math.sqrt(np.sum(np.square(xa)) * np.sum(np.square(xb))))
        if method == "euclidian_distance":                                        # IR:332-340
            return float(1 / (1 + math.sqrt(np.sum(np.square(np.subtract(a, b))))))
        if method == "manhattan_distance":                                        # IR:343-351 (square root included)
            return float(1 / (1 + math.sqrt(np.sum(np.abs(np.subtract(a, b))))))
        num = np.sum(np.multiply(a, b))
        a_sq, b_sq = np.sum(np.square(a)), np.sum(np.square(b))
        if method == "tanimoto_distance":                                         # IR:354-369
            return float(num / (a_sq + b_sq - num))
        if method == "dice_dist":                                                 # IR:372-389
            return float(2 * num / (a_sq + b_sq))
    raise ValueError(method)


def search(query: str, docs, method: str) -> np.ndarray:
    """Scores of every document in collection order (IR:466-470)."""
    a = represent(query, method)
    return np.array([score(method, a, represent(d, method)) for d in docs], dtype=np.float64)
