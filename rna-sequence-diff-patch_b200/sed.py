"""sed.py — the module surface of the reference's StringEditDistance.py on top of the CUDA engine.

Mirrors (file:line under /root/reference, SED = StringEditDistance.py):
  wagner_fisher      SED:133-224   values + tie edges come from rsd_matrix (fp64, on the GPU)
  create_paths       SED:228-271   host walk over the device-computed tie mask, reference BFS order
  generate_es        SED:274-334   generate_rev_es SED:338-369   generate_sequence_from_es SED:371-377
  patching           SED:380-457   cost SED:76-89   min_cost SED:92-128   Node/Edge SED:31-71
The functions take the cost dict explicitly; dropin/StringEditDistance.py binds them to the
module globals default_costs / user_costs exactly like the reference (SED:6-27).

Deviations (documented in DESIGN.md): dp is a lazy matrix view, not (m+1)(n+1) Python objects;
create_paths never deadlocks (the reference's bounded queue does, SED:237,265) and stops after
MAX_PATHS paths; nothing is printed at import."""
from __future__ import annotations

import numpy as np

from .encoding import SYMBOLS
from .engine import get_engine

MAX_PATHS = 1_000_000        # create_paths() cap; the number of co-optimal scripts is exponential
_CODE = {ch: i for i, ch in enumerate(SYMBOLS)}
_OPS = ("insert", "delete", "update")


# ------------------------------------------------------------------------------------------------
# symbols: reproduce the reference's lookup errors before anything reaches the GPU
# ------------------------------------------------------------------------------------------------
def _validate_and_encode(str1: str, str2: str, costs: dict):
    """The reference evaluates cost(str1[i-1], str2[j-1]) for every cell in row-major order
    (SED:185-187 -> SED:99 -> SED:79-87): equal-ignoring-case pairs are never looked up, every
    other pair indexes costs['update'][c1][c2].  Raise the KeyError it would raise first."""
    upd = costs["update"]
    if str1 and str2:
        seen = set()
        for c1 in str1:
            if c1 in seen:
                continue
            seen.add(c1)
            row = upd.get(c1)
            l1 = c1.lower()
            for c2 in str2:
                if c2.lower() == l1:
                    continue
                if row is None:
                    raise KeyError(c1)
                if c2 not in row:
                    raise KeyError(c2)
                # c1 row exists and holds c2: fine
            # no failing column for this row character
    def enc(s):
        out = np.empty(len(s), dtype=np.uint8)
        for k, ch in enumerate(s):
            c = _CODE.get(ch)
            if c is None:
                c = _CODE.get(ch.upper(), 15)
            out[k] = c
        return out
    return enc(str1), enc(str2)


# ------------------------------------------------------------------------------------------------
# Node / Edge / lazy dp matrix
# ------------------------------------------------------------------------------------------------
class Edge:
    """SED:31-39."""
    __slots__ = ("source", "destination", "operation")

    def __init__(self, source, destination, operation):
        self.source = source
        self.destination = destination
        self.operation = operation


class Node:
    """SED:42-71 — free-standing node (what callers may construct themselves)."""

    def __init__(self, i, j, value=0):
        self.i = i
        self.j = j
        self.value = value
        self.edges = []
        self.incoming_edges = []
        self.visited = False

    def add_neighbor(self, dest, operation):
        e = Edge(self, dest, operation)
        self.edges.append(e)
        dest.incoming_edges.append(e)

    def __repr__(self):
        return str(self.value)


class _CellNode:
    """Node view of one matrix cell; edges are materialised on demand from the tie mask."""
    __slots__ = ("_dp", "_r", "_c", "visited")

    def __init__(self, dp, r, c):
        self._dp, self._r, self._c = dp, r, c
        self.visited = False

    @property
    def i(self):            # string index, -1 on the border (SED:150,159,177,189)
        return self._r - 1

    @property
    def j(self):
        return self._c - 1

    @property
    def value(self):
        return self._dp._value(self._r, self._c)

    @property
    def incoming_edges(self):          # order insert, delete, update (SED:192-220)
        dp, r, c = self._dp, self._r, self._c
        mk = int(dp.mask[r, c])
        out = []
        if mk & 1: out.append(Edge(dp.node(r, c - 1), self, "insert"))
        if mk & 2: out.append(Edge(dp.node(r - 1, c), self, "delete"))
        if mk & 4: out.append(Edge(dp.node(r - 1, c - 1), self, "update"))
        return out

    @property
    def edges(self):
        dp, r, c = self._dp, self._r, self._c
        m, n = dp.m, dp.n
        ins = Edge(self, dp.node(r, c + 1), "insert") if c < n and dp.mask[r, c + 1] & 1 else None
        dele = Edge(self, dp.node(r + 1, c), "delete") if r < m and dp.mask[r + 1, c] & 2 else None
        upd = Edge(self, dp.node(r + 1, c + 1), "update") if r < m and c < n and dp.mask[r + 1, c + 1] & 4 else None
        # column 0 gets its delete edge in the border loop, before the interior loops (SED:167-182)
        order = (dele, ins, upd) if (c == 0 and r > 0) else (ins, dele, upd)
        return [e for e in order if e is not None]

    def __repr__(self):
        return str(self.value)

    def __eq__(self, other):
        return isinstance(other, _CellNode) and other._dp is self._dp and other._r == self._r and other._c == self._c

    def __hash__(self):
        return hash((id(self._dp), self._r, self._c))


class _Row:
    __slots__ = ("_dp", "_r")

    def __init__(self, dp, r):
        self._dp, self._r = dp, r

    def __len__(self):
        return self._dp.n + 1

    def __getitem__(self, c):
        if isinstance(c, slice):
            return [self._dp.node(self._r, k) for k in range(*c.indices(self._dp.n + 1))]
        n1 = self._dp.n + 1
        if c < 0:
            c += n1
        if not 0 <= c < n1:
            raise IndexError("list index out of range")
        return self._dp.node(self._r, c)

    def __iter__(self):
        return (self._dp.node(self._r, c) for c in range(self._dp.n + 1))

    def __repr__(self):
        return "[" + ", ".join(str(self._dp._value(self._r, c)) for c in range(self._dp.n + 1)) + "]"


class DPMatrix:
    """What wagnerFisher returns: indexable like list[list[Node]] (len(dp), len(dp[0]),
    dp[i][j].value — gui.py:364-379, IR:439) over device-computed fp64 values + tie masks."""

    def __init__(self, values: np.ndarray, mask: np.ndarray, str1: str, str2: str, costs: dict):
        self.values, self.mask = values, mask
        self.m, self.n = values.shape[0] - 1, values.shape[1] - 1
        self.str1, self.str2 = str1, str2
        self._costs = {"insert": costs["insert"], "delete": costs["delete"], "update": costs["update"]}
        self._nodes = {}
        self._isint = None

    def __len__(self):
        return self.m + 1

    def __getitem__(self, r):
        if isinstance(r, slice):
            return [_Row(self, k) for k in range(*r.indices(self.m + 1))]
        if r < 0:
            r += self.m + 1
        if not 0 <= r <= self.m:
            raise IndexError("list index out of range")
        return _Row(self, r)

    def __iter__(self):
        return (_Row(self, r) for r in range(self.m + 1))

    def __repr__(self):
        return "[" + ", ".join(repr(_Row(self, r)) for r in range(self.m + 1)) + "]"

    def node(self, r, c):
        key = (r, c)
        nd = self._nodes.get(key)
        if nd is None:
            nd = self._nodes[key] = _CellNode(self, r, c)
        return nd

    # -- Python number typing of the reference: int only where int + int was the first minimum --
    def _int_mask(self):
        if self._isint is not None:
            return self._isint
        m, n = self.m, self.n
        isint = np.zeros((m + 1, n + 1), dtype=bool)
        ins, dele, upd = self._costs["insert"], self._costs["delete"], self._costs["update"]
        ii = isinstance(ins, int) and not isinstance(ins, bool)
        di = isinstance(dele, int) and not isinstance(dele, bool)
        a, b = self.str1, self.str2

        def sub_is_int(r, c):
            c1, c2 = a[r - 1], b[c - 1]
            if c1.lower() == c2.lower():
                return True
            v = upd[c1][c2]
            return isinstance(v, int) and not isinstance(v, bool)

        isint[0, 0] = True
        if not ii and not di:
            # borders are float (j*ins, i*del), so an int can only travel down the main diagonal
            for k in range(1, min(m, n) + 1):
                if (self.mask[k, k] & 3) == 0 and sub_is_int(k, k):
                    isint[k, k] = True
                else:
                    break
        else:
            if ii: isint[0, 1:] = True
            if di: isint[1:, 0] = True
            for r in range(1, m + 1):
                for c in range(1, n + 1):
                    mk = int(self.mask[r, c])
                    if mk & 1: isint[r, c] = isint[r, c - 1] and ii
                    elif mk & 2: isint[r, c] = isint[r - 1, c] and di
                    else: isint[r, c] = isint[r - 1, c - 1] and sub_is_int(r, c)
        self._isint = isint
        return isint

    def _value(self, r, c):
        v = float(self.values[r, c])
        if (r == c or isinstance(self._costs["insert"], int) or isinstance(self._costs["delete"], int)) \
                and self._int_mask()[r, c]:
            return int(v)
        return v


# ------------------------------------------------------------------------------------------------
# the engine entry points
# ------------------------------------------------------------------------------------------------
def wagner_fisher(str1: str, str2: str, costs: dict, engine=None) -> DPMatrix:
    """SED:133-224.  str1 = source (rows), str2 = destination (columns)."""
    ca, cb = _validate_and_encode(str1, str2, costs)
    eng = engine or get_engine()
    eng.set_costs(costs)
    values, mask = eng.matrix(ca, cb)
    return DPMatrix(values, mask, str1, str2, costs)


def distance(str1: str, str2: str, costs: dict, engine=None) -> float:
    """dp[-1][-1].value without materialising the matrix (IR:437-439) — one-pair batch."""
    from .encoding import pack
    ca, cb = _validate_and_encode(str1, str2, costs)
    eng = engine or get_engine()
    eng.set_costs(costs)
    if max(len(ca), len(cb)) >= LONG_PAIR_MIN:                   # one long pair: the panel wavefront uses the whole GPU
        return float(eng.long_pair(ca, cb, want_script=False)["dist"])
    off_a = np.array([0, len(ca)], np.int64); off_b = np.array([0, len(cb)], np.int64)
    bits = 4 if (ca.size and ca.max() > 3) or (cb.size and cb.max() > 3) else 2
    d = eng.distance_batch(pack((ca, off_a), bits=bits), pack((cb, off_b), bits=bits))
    return float(d[0])


LONG_PAIR_MIN = 8192        # from this length on a single pair runs on the panel-wavefront kernels (rsd_long_pair)


def edit_script(str1: str, str2: str, costs: dict, engine=None):
    """generate_es(create_paths(wagnerFisher(str1, str2))[0], str1, str2) (SED:133-334) for a pair of ANY length, without
    the matrix: the canonical script straight from the device (rsd_script_batch, or rsd_long_pair from LONG_PAIR_MIN
    symbols on — including pairs whose direction matrix exceeds the device, which run in row blocks).  Empty strings
    raise IndexError like the reference (SED:278)."""
    from .encoding import pack
    if not str1 or not str2:
        raise IndexError("string index out of range")            # what generate_es does on a border-only path
    ca, cb = _validate_and_encode(str1, str2, costs)
    eng = engine or get_engine()
    eng.set_costs(costs)
    if max(len(ca), len(cb)) >= LONG_PAIR_MIN:
        res = eng.long_pair(ca, cb)
        return es_from_packed(res["op"], res["oi"], res["oj"], str1, str2)
    bits = 4 if ca.max() > 3 or cb.max() > 3 else 2
    off_a = np.array([0, len(ca)], np.int64); off_b = np.array([0, len(cb)], np.int64)
    res = eng.script_batch(pack((ca, off_a), bits=bits), pack((cb, off_b), bits=bits))
    k = int(res["n_ops"][0])
    return es_from_packed(res["op"][0, :k], res["oi"][0, :k], res["oj"][0, :k], str1, str2)


def _reach_lengths(mask: np.ndarray):
    """reach[r][c] = bitset of path lengths (edges) from (0,0) to (r,c) inside the tie DAG."""
    m1, n1 = mask.shape
    reach = [[0] * n1 for _ in range(m1)]
    reach[0][0] = 1
    for r in range(m1):
        row, mrow = reach[r], mask[r]
        up = reach[r - 1] if r else None
        for c in range(n1):
            if r == 0 and c == 0:
                continue
            mk = int(mrow[c]); v = 0
            if mk & 1: v |= row[c - 1]
            if mk & 2: v |= up[c]
            if mk & 4: v |= up[c - 1]
            row[c] = v << 1
    return reach


def iter_paths(dp: DPMatrix):
    """Every co-optimal path in the order of the reference's BFS (SED:244-271), lazily, as lists
    of matrix cells from (0,0) to (m,n).  BFS dequeue order == by length, then lexicographic in the
    predecessor choice (insert, delete, update) read from the sink — enumerated here with a
    length-bounded DFS so nothing exponential is ever held in memory."""
    mask = dp.mask
    m, n = dp.m, dp.n
    reach = _reach_lengths(mask)
    lengths = reach[m][n]
    L = 0
    while lengths >> L:
        if (lengths >> L) & 1:
            # DFS from the sink for paths of exactly L edges; stack holds (r, c, next choice)
            path = [(m, n)]
            choice = [0]
            while path:
                r, c = path[-1]
                depth = len(path) - 1
                if r == 0 and c == 0:
                    yield path[::-1]
                    path.pop(); choice.pop()
                    continue
                k = choice[-1]
                mk = int(mask[r, c])
                advanced = False
                while k < 3:
                    bit = 1 << k
                    k += 1
                    if not mk & bit:
                        continue
                    pr, pc = (r, c - 1) if bit == 1 else ((r - 1, c) if bit == 2 else (r - 1, c - 1))
                    if (reach[pr][pc] >> (L - depth - 1)) & 1:
                        choice[-1] = k
                        path.append((pr, pc)); choice.append(0)
                        advanced = True
                        break
                if not advanced:
                    path.pop(); choice.pop()
        L += 1


def create_paths(dp: DPMatrix, limit: int | None = None):
    """SED:228-271 -> list of paths, each a list of nodes (.i/.j string indices, -1 on borders),
    origin first (SED:270).  Stops after `limit` (default MAX_PATHS) paths instead of hanging."""
    cap = MAX_PATHS if limit is None else limit
    out = []
    for cells in iter_paths(dp):
        out.append([dp.node(r, c) for r, c in cells])
        if len(out) >= cap:
            break
    return out


def canonical_path(dp: DPMatrix):
    """create_paths(dp)[0] without enumerating."""
    for cells in iter_paths(dp):
        return [dp.node(r, c) for r, c in cells]
    return []


def generate_es(path, str1: str, str2: str):
    """SED:274-334.  One op dict per edge; matches are 'update' with equal characters; indices are
    node indices (cell - 1), and -1 wraps to the last character like the reference's str1[next.i]."""
    nxt = path[1]                       # IndexError on a one-node path, like SED:278
    es = []
    for k in range(1, len(path)):
        cur, nxt = path[k - 1], path[k]
        di, dj = nxt.i - cur.i, nxt.j - cur.j
        op = "update" if (di == 1 and dj == 1) else ("delete" if di == 1 else "insert")
        es.append({"operation": op,
                   "source": {"character": str1[nxt.i], "index": nxt.i},
                   "destination": {"character": str2[nxt.j], "index": nxt.j}})
    return es


def es_from_packed(op, oi, oj, str1: str, str2: str):
    """Packed script of rsd_script_batch (ops + entered cells) -> the reference's list of dicts."""
    return [{"operation": _OPS[int(o)],
             "source": {"character": str1[int(i) - 1], "index": int(i) - 1},
             "destination": {"character": str2[int(j) - 1], "index": int(j) - 1}}
            for o, i, j in zip(op, oi, oj)]


def generate_rev_es(es):
    """SED:338-369 — insert<->delete, update swapped; inner dicts are shared, not copied."""
    out = []
    for e in es:
        kind = e["operation"]
        if kind == "insert":
            rev = {"operation": "delete", "source": e["destination"], "destination": e["source"]}
        elif kind == "delete":
            rev = {"operation": "insert",
                   "source": {"index": e["destination"]["index"] - 1, "character": e["destination"]["character"]},
                   "destination": e["source"]}
        elif kind == "update":
            rev = {"operation": "update", "source": e["destination"], "destination": e["source"]}
        else:                               # the reference would reuse stale locals / NameError
            raise ValueError(f"unknown operation {kind!r}")
        out.append(rev)
    return out


def generate_sequence_from_es(es):
    """SED:371-377."""
    return "".join(e["source"]["character"] for e in es if e["operation"] != "insert")


def patching(es, str1: str):
    """SED:380-457 — (error_code, patched).  Sequential semantics so hand-edited scripts behave
    like the reference; the batched GPU path (Engine.patch_batch) uses the closed form that holds
    for generated scripts."""
    expected = generate_sequence_from_es(es)
    if str1 == expected:
        code = 0
    elif len(str1) >= len(expected):
        code = 1
    else:
        return (-1, "")
    s = str1
    shift = 0                                   # (#inserts - #deletes) applied so far
    for e in es:
        kind = e["operation"]
        if kind == "insert":
            at = e["destination"]["index"]
            s = s[:at] + e["destination"]["character"] + s[at:]
            shift += 1
        else:
            at = e["source"]["index"] + shift
            if kind == "update":
                s = s[:at] + e["destination"]["character"] + s[at + 1:]
            elif kind == "delete":
                s = s[0:at] + s[at + 1:]
                shift -= 1
    return (code, s)


def cost(char1: str, char2: str, costs: dict):
    """SED:76-89."""
    if char1.lower() == char2.lower():
        return 0
    return costs["update"][char1][char2]


def min_cost(dp, i, j, str1, str2, costs: dict):
    """SED:92-128 on any dp exposing dp[i][j].value."""
    cands = [dp[i][j - 1].value + costs["insert"],
             dp[i - 1][j].value + costs["delete"],
             dp[i - 1][j - 1].value + cost(str1[i - 1], str2[j - 1], costs)]
    val = min(cands)
    preds = ((i, j - 1, "insert"), (i - 1, j, "delete"), (i - 1, j - 1, "update"))
    return val, [p if c == val else None for p, c in zip(preds, cands)]


def format_edit_script(es) -> str:
    """gui.py:72-90."""
    parts = []
    for e in es:
        kind = e["operation"]
        if kind == "update" and e["source"]["character"] == e["destination"]["character"]:
            continue
        if kind == "insert":
            parts.append(f'Ins({e["source"]["index"]},{e["destination"]["character"]})')
        elif kind == "delete":
            parts.append(f'Del({e["source"]["index"]})')
        else:
            parts.append(f'Upd({e["source"]["index"]},{e["destination"]["character"]})')
    return "[" + ",".join(parts) + "]"
