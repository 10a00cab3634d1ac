"""oracle/oracle.py — CPU oracle for the weighted Wagner–Fischer path (TEST INFRASTRUCTURE ONLY).

Nothing in the product package may import this module.  Allowed importers: tests/,
__graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).

Two layers:
  * ctypes bindings to oracle/liborc.so (wf_oracle.c): fp64 distance / matrix+tie-mask /
    canonical script / closed-form patch / search top-k, single and batched (pthreads).
  * pure-Python restatements of the object-level functions of the reference
    (create_paths BFS order, generate_es, generate_rev_es, patching, format_edit_script,
    Python int/float value typing) — loops, so small cases only.

Parity status: PINNED by tests/test_oracle_golden.py against tests/golden/ref_golden.json, which
tests/golden/make_golden.py produced by running the unmodified reference in the build container.
Citations: SED = /root/reference/StringEditDistance.py, IR = /root/reference/IRMethods.py.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SYMBOLS = "AGCUYRWSKMDVHBN"          # IR:13 — row/column order of costs.json
_CODE = {ch: i for i, ch in enumerate(SYMBOLS)}
OPS = ("insert", "delete", "update")  # candidate order SED:103


class OrcCosts(C.Structure):
    _fields_ = [("ins", C.c_double), ("del_", C.c_double), ("sub", (C.c_double * 16) * 16)]


def build(force: bool = False) -> str:
    """Compile wf_oracle.c -> oracle/liborc.so (gcc, no OpenMP dependency)."""
    so = os.path.join(HERE, "liborc.so")
    src = os.path.join(HERE, "wf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math",
                               "-shared", "-o", so, src])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, i64p, i32p, f64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_double))
        cp = C.POINTER(OrcCosts)
        L.orc_distance.restype = C.c_double
        L.orc_distance.argtypes = [u8p, C.c_int, u8p, C.c_int, cp]
        L.orc_matrix.restype = None
        L.orc_matrix.argtypes = [u8p, C.c_int, u8p, C.c_int, cp, f64p, u8p]
        L.orc_canonical_script.restype = C.c_int
        L.orc_canonical_script.argtypes = [u8p, C.c_int, u8p, C.c_int, cp, u8p, i32p, i32p, f64p]
        L.orc_distance_batch.restype = None
        L.orc_distance_batch.argtypes = [u8p, i64p, u8p, i64p, C.c_int64, cp, f64p, C.c_int]
        L.orc_script_batch.restype = None
        L.orc_script_batch.argtypes = [u8p, i64p, u8p, i64p, C.c_int64, cp, C.c_int64, u8p, i32p, i32p,
                                       i32p, f64p, C.c_int]
        L.orc_patch_closed.restype = C.c_int
        L.orc_patch_closed.argtypes = [u8p, i32p, i32p, C.c_int, u8p, C.c_int, u8p, C.c_int, u8p, C.c_int,
                                       u8p, C.POINTER(C.c_int)]
        L.orc_search_topk.restype = None
        L.orc_search_topk.argtypes = [u8p, C.c_int, u8p, i64p, C.c_int64, cp, C.c_int, i64p, f64p, f64p,
                                      C.c_int]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(arr, typ):
    return arr.ctypes.data_as(C.POINTER(typ))


def load_costs(path: str) -> dict:
    with open(path) as f:
        return json.load(f)


def costs_struct(costs: dict) -> OrcCosts:
    s = OrcCosts()
    s.ins = float(costs["insert"])
    s.del_ = float(costs["delete"])
    for a, row in costs["update"].items():
        for b, v in row.items():
            s.sub[_CODE[a]][_CODE[b]] = float(v)
    return s


def encode(s: str) -> np.ndarray:
    """Uppercase table symbols -> 0..14 (IR:13).  Anything else is a caller error here."""
    return np.array([_CODE[ch] for ch in s], dtype=np.uint8)


def decode(codes) -> str:
    return "".join(SYMBOLS[c] for c in codes)


def concat(seqs):
    """list[str] -> (codes uint8, offsets int64[n+1])."""
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    for i, s in enumerate(seqs):
        off[i + 1] = off[i] + len(s)
    codes = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
    for i, s in enumerate(seqs):
        if s:
            codes[off[i]:off[i + 1]] = encode(s)
    return codes, off


# ------------------------------------------------------------------ C-backed functions
def distance(a: str, b: str, costs: dict) -> float:
    ca, cb = encode(a) if a else np.zeros(1, np.uint8), encode(b) if b else np.zeros(1, np.uint8)
    cs = costs_struct(costs)
    return lib().orc_distance(_p(ca, C.c_uint8), len(a), _p(cb, C.c_uint8), len(b), C.byref(cs))


def matrix(a: str, b: str, costs: dict):
    ca, cb = encode(a) if a else np.zeros(1, np.uint8), encode(b) if b else np.zeros(1, np.uint8)
    cs = costs_struct(costs)
    D = np.zeros((len(a) + 1, len(b) + 1), dtype=np.float64)
    M = np.zeros((len(a) + 1, len(b) + 1), dtype=np.uint8)
    lib().orc_matrix(_p(ca, C.c_uint8), len(a), _p(cb, C.c_uint8), len(b), C.byref(cs),
                     _p(D, C.c_double), _p(M, C.c_uint8))
    return D, M


def canonical_script(a: str, b: str, costs: dict):
    """-> (ops uint8[k], i int32[k], j int32[k], distance); (i,j) = matrix cell each op enters."""
    ca, cb = encode(a) if a else np.zeros(1, np.uint8), encode(b) if b else np.zeros(1, np.uint8)
    cs = costs_struct(costs)
    cap = len(a) + len(b) + 1
    ops = np.zeros(cap, np.uint8); oi = np.zeros(cap, np.int32); oj = np.zeros(cap, np.int32)
    d = C.c_double()
    k = lib().orc_canonical_script(_p(ca, C.c_uint8), len(a), _p(cb, C.c_uint8), len(b), C.byref(cs),
                                   _p(ops, C.c_uint8), _p(oi, C.c_int32), _p(oj, C.c_int32), C.byref(d))
    assert k >= 0
    return ops[:k].copy(), oi[:k].copy(), oj[:k].copy(), d.value


def distance_batch(a_codes, a_off, b_codes, b_off, costs: dict, nthreads: int = 0) -> np.ndarray:
    n = len(a_off) - 1
    out = np.zeros(n, np.float64)
    cs = costs_struct(costs)
    lib().orc_distance_batch(_p(a_codes, C.c_uint8), _p(a_off, C.c_int64), _p(b_codes, C.c_uint8),
                             _p(b_off, C.c_int64), n, C.byref(cs), _p(out, C.c_double), nthreads)
    return out


def script_batch(a_codes, a_off, b_codes, b_off, costs: dict, nthreads: int = 0):
    n = len(a_off) - 1
    la = np.diff(a_off); lb = np.diff(b_off)
    max_ops = int((la + lb).max()) + 1 if n else 1
    ops = np.zeros((n, max_ops), np.uint8); oi = np.zeros((n, max_ops), np.int32)
    oj = np.zeros((n, max_ops), np.int32); cnt = np.zeros(n, np.int32); dist = np.zeros(n, np.float64)
    cs = costs_struct(costs)
    lib().orc_script_batch(_p(a_codes, C.c_uint8), _p(a_off, C.c_int64), _p(b_codes, C.c_uint8),
                           _p(b_off, C.c_int64), n, C.byref(cs), max_ops, _p(ops, C.c_uint8),
                           _p(oi, C.c_int32), _p(oj, C.c_int32), _p(cnt, C.c_int32), _p(dist, C.c_double),
                           nthreads)
    return ops, oi, oj, cnt, dist


def patch_closed_codes(ops, oi, oj, a_codes, b_codes, x_codes):
    out = np.zeros(len(x_codes) + len(ops) + 1, np.uint8)
    ol = C.c_int()
    ops = np.ascontiguousarray(ops, np.uint8); oi = np.ascontiguousarray(oi, np.int32)
    oj = np.ascontiguousarray(oj, np.int32)
    a_codes = np.ascontiguousarray(a_codes, np.uint8); b_codes = np.ascontiguousarray(b_codes, np.uint8)
    x_codes = np.ascontiguousarray(x_codes, np.uint8)
    code = lib().orc_patch_closed(_p(ops, C.c_uint8), _p(oi, C.c_int32), _p(oj, C.c_int32), len(ops),
                                  _p(a_codes, C.c_uint8), len(a_codes), _p(b_codes, C.c_uint8), len(b_codes),
                                  _p(x_codes, C.c_uint8), len(x_codes), _p(out, C.c_uint8), C.byref(ol))
    return code, out[:ol.value].copy()


def search_topk(query: str, db_codes, db_off, costs: dict, k: int, nthreads: int = 0, want_scores=False):
    n = len(db_off) - 1
    q = encode(query)
    idx = np.zeros(k, np.int64); sc = np.zeros(k, np.float64)
    allsc = np.zeros(n, np.float64) if want_scores else None
    cs = costs_struct(costs)
    lib().orc_search_topk(_p(q, C.c_uint8), len(query), _p(db_codes, C.c_uint8), _p(db_off, C.c_int64), n,
                          C.byref(cs), k, _p(idx, C.c_int64), _p(sc, C.c_double),
                          _p(allsc, C.c_double) if want_scores else None, nthreads)
    return (idx, sc, allsc) if want_scores else (idx, sc)


def num_threads() -> int:
    return lib().orc_num_threads()


# ------------------------------------------------------------------ pure-Python restatements
def py_matrix(a: str, b: str, costs: dict):
    """SED:133-224 with Python numbers, so int/float typing (and str()) match the reference.
    -> (values list[list], mask list[list]) ; mask bits: 1 INS, 2 DEL, 4 UPD."""
    m, n = len(a), len(b)
    ins, dele, upd = costs["insert"], costs["delete"], costs["update"]
    D = [[0] * (n + 1) for _ in range(m + 1)]
    M = [[0] * (n + 1) for _ in range(m + 1)]
    for j in range(1, n + 1):
        D[0][j] = j * ins; M[0][j] = 1                       # SED:159
    for i in range(1, m + 1):
        D[i][0] = i * dele; M[i][0] = 2                      # SED:177
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            c1, c2 = a[i - 1], b[j - 1]
            sub = 0 if c1.lower() == c2.lower() else upd[c1][c2]   # SED:79-87
            cands = [D[i][j - 1] + ins, D[i - 1][j] + dele, D[i - 1][j - 1] + sub]  # SED:95-103
            v = min(cands)                                          # SED:107
            D[i][j] = v
            M[i][j] = sum(1 << k for k, c in enumerate(cands) if c == v)  # SED:109
    return D, M


def all_paths(mask, cap=None):
    """SED:228-271 — every co-optimal path, in the reference's BFS dequeue order, each as a list of
    matrix cells (i,j) from (0,0) to (m,n).  `cap` bounds the number of *completed* paths."""
    m, n = len(mask) - 1, len(mask[0]) - 1
    q = deque([[(m, n)]])
    done = []
    while q:
        p = q.popleft()
        i, j = p[-1]
        if i == 0 and j == 0:
            done.append(p[::-1])
            if cap is not None and len(done) >= cap:
                break
            continue
        mk = mask[i][j]
        if mk & 1: q.append(p + [(i, j - 1)])
        if mk & 2: q.append(p + [(i - 1, j)])
        if mk & 4: q.append(p + [(i - 1, j - 1)])
    return done


def es_from_cells(cells, a: str, b: str):
    """SED:274-334 on a path given as matrix cells.  Node indices are (i-1, j-1) (SED:159,177,189);
    negative indices wrap like Python's (SED:302-323).  Raises IndexError on empty strings like the
    reference does (SED:278,302)."""
    if len(cells) < 2:
        raise IndexError("list index out of range")           # path[1] SED:278
    es = []
    for (pi, pj), (ci, cj) in zip(cells[:-1], cells[1:]):
        op = "update" if (ci == pi + 1 and cj == pj + 1) else ("delete" if ci == pi + 1 else "insert")
        es.append({"operation": op,
                   "source": {"character": a[ci - 1], "index": ci - 1},
                   "destination": {"character": b[cj - 1], "index": cj - 1}})
    return es


def es_from_ops(ops, oi, oj, a: str, b: str):
    return [{"operation": OPS[int(o)],
             "source": {"character": a[int(i) - 1], "index": int(i) - 1},
             "destination": {"character": b[int(j) - 1], "index": int(j) - 1}}
            for o, i, j in zip(ops, oi, oj)]


def rev_es(es):
    """SED:338-369."""
    out = []
    for e in es:
        op = e["operation"]
        if op == "insert":
            out.append({"operation": "delete", "source": e["destination"], "destination": e["source"]})
        elif op == "delete":
            out.append({"operation": "insert",
                        "source": {"index": e["destination"]["index"] - 1,
                                   "character": e["destination"]["character"]},
                        "destination": e["source"]})
        else:
            out.append({"operation": "update", "source": e["destination"], "destination": e["source"]})
    return out


def seq_from_es(es):
    """SED:371-377."""
    return "".join(op["source"]["character"] for op in es if op["operation"] != "insert")


def patch_sequential(es, s: str):
    """SED:380-457, statement for statement."""
    orig = seq_from_es(es)
    if s == orig:
        code = 0
    elif len(s) >= len(orig):
        code = 1
    else:
        return (-1, "")
    n_del = 0
    n_ins = 0
    for e in es:
        op = e["operation"]
        si = e["source"]["index"]
        di = e["destination"]["index"]
        si = si + n_del + n_ins if op != "insert" else di     # SED:422-427
        if op == "update":
            s = s[:si] + e["destination"]["character"] + s[si + 1:]
        if op == "delete":
            s = s[0:si:] + s[si + 1::]
            n_del -= 1
        if op == "insert":
            s = s[:si] + e["destination"]["character"] + s[si:]
            n_ins += 1
    return (code, s)


def patch_closed(es, s: str):
    """Closed form valid for scripts produced by generate_es / generate_rev_es."""
    orig = seq_from_es(es)
    if s == orig:
        code = 0
    elif len(s) >= len(orig):
        code = 1
    else:
        return (-1, "")
    out = "".join(e["destination"]["character"] for e in es if e["operation"] != "delete")
    return (code, out + s[len(orig):])


def format_es(es) -> str:
    """gui.py:72-90."""
    parts = []
    for op in es:
        o = op["operation"]
        if o == "update" and op["source"]["character"] == op["destination"]["character"]:
            continue
        if o == "insert":
            parts.append(f'Ins({op["source"]["index"]},{op["destination"]["character"]})')
        elif o == "delete":
            parts.append(f'Del({op["source"]["index"]})')
        else:
            parts.append(f'Upd({op["source"]["index"]},{op["destination"]["character"]})')
    return "[" + ",".join(parts) + "]"


def wf_score(a: str, b: str, costs: dict) -> float:
    """IR:435-440."""
    return 1 / (1 + distance(a, b, costs))


def topk_stable(scores, k):
    """performance.py:12-15 on a list of (item, score)."""
    return sorted(scores, key=lambda t: t[1], reverse=True)[:k]
