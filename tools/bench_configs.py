#!/usr/bin/env python
"""Measure BASELINE configs 3, 4 and 5 (bench.py covers config 2, the headline).  One JSON line per
config on stdout; used for profiles/ and DESIGN.md, not by the driver.
    python tools/bench_configs.py [--c3-pairs N] [--c5-records N] [--which c3,c4,c5[,c5i]]   (c5i: C5 with 1 % IUPAC records, fp64)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402

G.build()
import rna_sequence_diff_patch_b200 as R  # noqa: E402

DROPIN = os.path.join(ROOT, "rna-sequence-diff-patch_b200", "dropin")
DEFAULT = __import__('rna_sequence_diff_patch_b200.cost_tables', fromlist=['x']).default_costs()


def mutate_batch(rng, codes, off, alpha=4, p_sub=0.05, p_ins=0.025, p_del=0.025, lo=1000, hi=2000):
    """vectorised per-base mutation of a concatenated batch, each result clipped into [lo, hi]"""
    out_c, out_l = [], []
    for p in range(len(off) - 1):
        a = codes[off[p]:off[p + 1]]
        r = rng.random(a.shape[0])
        sub = rng.integers(0, alpha, size=a.shape[0], dtype=np.uint8)
        x = np.where(r < p_sub, sub, a)
        keep = ~((r >= p_sub) & (r < p_sub + p_del))
        reps = np.where(r > 1 - p_ins, 2, 1) * keep
        y = np.repeat(x, reps)
        dup = np.repeat(np.arange(a.shape[0]), reps)
        second = np.concatenate([[False], dup[1:] == dup[:-1]])
        y[second] = rng.integers(0, alpha, size=int(second.sum()), dtype=np.uint8)
        if y.shape[0] > hi: y = y[:hi]
        if y.shape[0] < lo: y = np.concatenate([y, rng.integers(0, alpha, size=lo - y.shape[0], dtype=np.uint8)])
        out_c.append(y); out_l.append(y.shape[0])
    o = np.zeros(len(out_l) + 1, np.int64); np.cumsum(out_l, out=o[1:])
    return np.concatenate(out_c), o


def c3(eng, n_pairs, reps=3):
    rng = np.random.default_rng(20260003)
    la = rng.integers(1000, 2001, size=n_pairs)
    oa = np.zeros(n_pairs + 1, np.int64); np.cumsum(la, out=oa[1:])
    ca = rng.integers(0, 4, size=int(oa[-1]), dtype=np.uint8)
    cb, ob = mutate_batch(rng, ca, oa)
    indep = np.arange(n_pairs) % 10 == 0
    for p in np.nonzero(indep)[0]:
        cb[ob[p]:ob[p + 1]] = rng.integers(0, 4, size=int(ob[p + 1] - ob[p]), dtype=np.uint8)
    A, B = R.pack((ca, oa)), R.pack((cb, ob))
    cells = float((np.diff(oa) * np.diff(ob)).sum())
    eng.set_costs(DEFAULT); eng.set_timing(True)
    import ctypes as C
    from rna_sequence_diff_patch_b200 import _lib
    lib = R.load_library()
    max_ops = int((A.len.astype(np.int64) + B.len).max())
    def pinned(shape, dtype):                     # page-locked output buffers (rsd_host_alloc), like a serving caller would use
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = C.c_void_p(); _lib.check(lib.rsd_host_alloc(C.byref(ptr), n))
        return np.frombuffer((C.c_uint8 * n).from_address(ptr.value), dtype=dtype).reshape(shape)
    for P in (A, B):                              # page-locked inputs as well
        for name in ("words", "start", "len"):
            src = getattr(P, name); dst = pinned(src.shape, src.dtype); dst[...] = src; setattr(P, name, dst)
    op = pinned((n_pairs, max_ops), np.uint8); n_ops = pinned((n_pairs,), np.int32); dist = pinned((n_pairs,), np.float64)
    ok = pinned((n_pairs,), np.uint8); mode = C.c_int()
    ts, kms = [], []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        _lib.check(lib.rsd_script_patch_check_batch(
            eng.ctx, _lib.ptr(A.words, C.c_uint32), _lib.ptr(A.start, C.c_int64), _lib.ptr(A.len, C.c_int32), A.words.shape[0],
            _lib.ptr(B.words, C.c_uint32), _lib.ptr(B.start, C.c_int64), _lib.ptr(B.len, C.c_int32), B.words.shape[0],
            n_pairs, A.bits, A.symmask | B.symmask, 0, max_ops, _lib.ptr(op, C.c_uint8), None, None,
            _lib.ptr(n_ops, C.c_int32), _lib.ptr(dist, C.c_double), _lib.ptr(ok, C.c_uint8), C.byref(mode)))
        if r:
            ts.append(time.perf_counter() - t0); kms.append(eng.last_kernel_ms())
    assert ok.all()
    t, k = float(np.mean(ts)), float(np.mean(kms)) * 1e-3
    # HBM streams of the script path: 2-bit direction codes written once by the forward kernel (16 rows per
    # 32-bit word, columns padded to the 32-column strips), packed inputs read, op bytes written
    la_, lb_ = np.diff(oa), np.diff(ob)
    dir_bytes = float((((la_ + 15) // 16) * (((lb_ + 31) // 32) * 32) * 4).sum())
    in_bytes = float(A.words.nbytes + B.words.nbytes + 2 * n_pairs * 12)
    hbm = {"direction_bytes_written": dir_bytes, "packed_input_bytes": in_bytes, "op_bytes_written": float(n_ops.sum()),
           "achieved_gbs_over_device_time": (dir_bytes + in_bytes + float(n_ops.sum())) / k * 1e-9, "peak_gbs": 6552.6}
    return {"config": "C3", "hbm_streams": hbm, "pairs": n_pairs, "cells": cells, "mode": mode.value,
            "e2e_pairs_per_s": n_pairs / t, "e2e_gcups": cells / t * 1e-9, "e2e_s": t,
            "device_pairs_per_s": n_pairs / k, "device_gcups": cells / k * 1e-9, "device_s": k,
            "roundtrip_ok": bool(ok.all()), "output": "op bytes + n_ops + dist + ok (oi/oj derivable by prefix sum), pinned host buffers"}


def c4(eng, L=50000, reps=3):
    rng = np.random.default_rng(20260004)
    a = rng.integers(0, 4, size=L, dtype=np.uint8)
    cb, ob = mutate_batch(rng, a, np.array([0, L]), lo=1, hi=10 ** 9)
    b = cb
    eng.set_costs(DEFAULT); eng.set_timing(True)
    ts, kms = [], []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        res = eng.long_pair(a, b)
        if r:
            ts.append(time.perf_counter() - t0); kms.append(eng.last_kernel_ms())
    cells = float(L) * b.shape[0]
    t, k = float(np.mean(ts)), float(np.mean(kms)) * 1e-3
    fwd = []
    for r in range(reps + 1):
        eng.long_pair(a, b, want_script=False)
        if r:
            fwd.append(eng.last_kernel_ms())
    kf = float(np.mean(fwd)) * 1e-3
    return {"config": "C4", "forward_only_device_s": kf, "forward_only_gcups": cells / kf * 1e-9, "m": L, "n": int(b.shape[0]), "cells": cells, "mode": res["mode"], "dist": res["dist"],
            "n_ops": int(res["op"].shape[0]), "e2e_gcups": cells / t * 1e-9, "e2e_s": t,
            "device_gcups": cells / k * 1e-9, "device_s": k}


def c5(eng, n_rec, nq=64, k=10, reps=3, iupac=False):
    rng = np.random.default_rng(20260005)
    lens = rng.integers(24, 32, size=n_rec)
    off = np.zeros(n_rec + 1, np.int64); np.cumsum(lens, out=off[1:])
    codes = rng.integers(0, 4, size=int(off[-1]), dtype=np.uint8)
    codes[rng.random(codes.shape[0]) < 1e-3] = 14
    if iupac:                                           # second run, 1 % of the records drawn from all 15 symbols
        full = np.repeat(rng.random(n_rec) < 0.01, lens)   #   -> non-dyadic default costs (0.66 / 0.83) -> fp64 reference-order kernels
        codes[full] = rng.integers(0, 15, size=int(full.sum()), dtype=np.uint8)
    qs, qo = [], [0]
    for r in rng.integers(0, n_rec, size=nq):
        s = codes[off[r]:off[r + 1]].copy()
        hit = rng.random(s.shape[0]) < 0.1
        s[hit] = rng.integers(0, 4, size=int(hit.sum()), dtype=np.uint8)
        qs.append(s); qo.append(qo[-1] + s.shape[0])
    Q = R.pack((np.concatenate(qs), np.array(qo, np.int64)), bits=4)
    db = R.pack((codes, off), bits=4)
    eng.set_costs(DEFAULT); eng.set_timing(True)
    eng.db_load(db)
    cells = float(sum(len(q) for q in qs)) * 0 + float(np.add.outer(np.array([len(q) for q in qs]), np.zeros(1)).sum() * 0)
    cells = float(np.array([len(q) for q in qs], np.float64).sum() * lens.astype(np.float64).sum())
    ts, kms = [], []
    for r in range(reps + 1):
        t0 = time.perf_counter()
        idx, sc = eng.db_search_topk(Q, k)
        if r:
            ts.append(time.perf_counter() - t0); kms.append(eng.last_kernel_ms())
    # the other scorers of search_collection on the same database (8f rank 4): one query, all scores + top-k on device
    sim = {}
    if iupac:
        eng.db_free()
        t, kk = float(np.mean(ts)), float(np.mean(kms)) * 1e-3
        return {"config": "C5-iupac", "records": n_rec, "queries": nq, "k": k, "cells": cells, "mode": eng.last_mode,
                "e2e_gcups": cells / t * 1e-9, "e2e_s": t, "device_gcups": cells / kk * 1e-9, "device_s": kk}
    db_bytes = float(db.words.nbytes + n_rec * (8 + 4 + 8 + 8))           # words + start + len + perm read, score written
    for method in ("set_jaccard_similarity", "multi_dice_similarity", "cosine", "pearson"):
        ms = []
        for r in range(3):
            eng.db_similarity(qs[0], method, k=k, want_scores=False)
            if r:
                ms.append(eng.last_kernel_ms())
        m = float(np.mean(ms))
        sim[method] = {"ms_per_query": m, "records_per_s": n_rec / (m * 1e-3), "hbm_gbs_algorithmic": db_bytes / (m * 1e-3) * 1e-9}
    eng.db_free()
    t, kk = float(np.mean(ts)), float(np.mean(kms)) * 1e-3
    return {"config": "C5", "similarity_search": sim, "records": n_rec, "queries": nq, "k": k, "cells": cells, "mode": eng.last_mode,
            "e2e_gcups": cells / t * 1e-9, "e2e_s": t, "device_gcups": cells / kk * 1e-9, "device_s": kk,
            "top1_scores_head": sc[:3, 0].tolist()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c3,c4,c5")
    ap.add_argument("--c3-pairs", type=int, default=100000)
    ap.add_argument("--c5-records", type=int, default=10_000_000)
    ap.add_argument("--c5-iupac-queries", type=int, default=8)
    args = ap.parse_args()
    eng = R.Engine(0)
    for w in args.which.split(","):
        t0 = time.perf_counter()
        res = {"c3": lambda: c3(eng, args.c3_pairs), "c4": lambda: c4(eng), "c5": lambda: c5(eng, args.c5_records),
               "c5i": lambda: c5(eng, args.c5_records, nq=args.c5_iupac_queries, reps=2, iupac=True)}[w]()
        res["wall_incl_datagen_s"] = time.perf_counter() - t0
        print(json.dumps(res), flush=True)
