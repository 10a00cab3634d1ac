"""Synthetic inputs shared by the GPU tests and the golden generators (B = A mutated per base)."""
import numpy as np


def mutate_codes(rng, a, alpha=4, p_sub=0.05, p_ins=0.025, p_del=0.025):
    r = rng.random(a.shape[0])
    keep = r >= p_del
    out = []
    sub = rng.integers(0, alpha, size=a.shape[0], dtype=np.uint8)
    ins = rng.integers(0, alpha, size=a.shape[0], dtype=np.uint8)
    for k in range(a.shape[0]):
        if not keep[k]:
            continue
        out.append(sub[k] if r[k] < p_del + p_sub else a[k])
        if r[k] > 1 - p_ins:
            out.append(ins[k])
    return np.array(out, dtype=np.uint8)


def c4_pair(seed=20260004, m=50000):
    """The BASELINE config 4 pair: 50 kb of ACGU and its mutated copy."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 4, size=m, dtype=np.uint8)
    return a, mutate_codes(rng, a)
