#!/usr/bin/env python
"""Generate tests/golden/ref_golden.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What it does: chdir to /root/reference (the reference opens costs.json / user_costs.json relative
to the CWD, SED:6-18), import StringEditDistance, IRMethods and import_xml with stdout captured
(SED:463-471 prints a self-test at import), replace StringEditDistance.Queue by an unbounded queue
(the bounded one deadlocks, SED:237,265), then record inputs and the reference's outputs.
Nothing from the reference is copied into the repository except these input/output vectors.
"""
import contextlib
import io
import json
import os
import queue
import random
import sys
from operator import itemgetter

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.json")
NUC = ['A', 'G', 'C', 'U', 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N']  # IR:13

sys.dont_write_bytecode = True
os.chdir(REF)
sys.path.insert(0, REF)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    import StringEditDistance as S
    import IRMethods as IR
    from import_xml import import_xml
IMPORT_STDOUT = buf.getvalue()


class _Unbounded(queue.Queue):
    def __init__(self, maxsize=0):
        super().__init__(0)


S.Queue = _Unbounded


def tie_mask(dp):
    out = []
    for row in dp:
        r = []
        for node in row:
            mk = 0
            for e in node.incoming_edges:
                mk |= {"insert": 1, "delete": 2, "update": 4}[e.operation]
            r.append(mk)
        out.append(r)
    return out


def count_paths(mask):
    m, n = len(mask), len(mask[0])
    cnt = [[0] * n for _ in range(m)]
    cnt[0][0] = 1
    for i in range(m):
        for j in range(n):
            if i == 0 and j == 0:
                continue
            c = 0
            if mask[i][j] & 1: c += cnt[i][j - 1]
            if mask[i][j] & 2: c += cnt[i - 1][j]
            if mask[i][j] & 4: c += cnt[i - 1][j - 1]
            cnt[i][j] = c
    return cnt[m - 1][n - 1]


def full_case(a, b, user, want_all_paths=True, path_cap=4000, keep_paths=40, want_dp=True):
    """Run the whole L1 pipeline of the reference on one pair."""
    dp = S.wagnerFisher(a, b, user)
    rec = {"a": a, "b": b, "user": user,
           "distance": dp[len(dp) - 1][len(dp[0]) - 1].value,
           "distance_repr": repr(dp[len(dp) - 1][len(dp[0]) - 1].value)}
    mask = tie_mask(dp)
    if want_dp:
        rec["dp_repr"] = [[repr(x.value) for x in row] for row in dp]
        rec["mask"] = mask
    npaths = count_paths(mask)
    rec["n_paths"] = npaths
    if want_all_paths and npaths <= path_cap and a and b:
        paths = S.create_paths(dp)
        assert len(paths) == npaths, (len(paths), npaths)
        rec["paths_cells"] = [[(nd.i + 1, nd.j + 1) for nd in p] for p in paths[:keep_paths]]
        ess = [S.generate_es(p, a, b) for p in paths[:keep_paths]]
        rec["es"] = ess[:3]
        rec["es_fmt"] = [fmt(e) for e in ess]
        es0 = ess[0]
        rec["patch0"] = list(S.patching(es0, a))
        rev = S.generate_rev_es(es0)
        rec["rev0"] = rev
        rec["rev_patch0"] = list(S.patching(rev, b))
        rec["seq_from_es0"] = S.generate_sequence_from_es(es0)
    return rec


def fmt(es):  # gui.py:72-90 cannot be imported (PySide2); this is the golden *input* formatter only,
    # its expected strings G2-G5 are pinned in SURVEY Appendix B and re-checked in the tests.
    out = []
    for op in es:
        o = op['operation']
        if o == 'update' and op['source']['character'] == op['destination']['character']:
            continue
        if o == 'insert':
            out.append(f"Ins({op['source']['index']},{op['destination']['character']})")
        elif o == 'delete':
            out.append(f"Del({op['source']['index']})")
        else:
            out.append(f"Upd({op['source']['index']},{op['destination']['character']})")
    return '[' + ','.join(out) + ']'


G = {"import_stdout": IMPORT_STDOUT,
     "default_costs": S.default_costs, "user_costs": S.user_costs}

# ---- G1: the import-time self test + patch error codes --------------------------------------
g1 = full_case('AGRGA', 'AGGGAA', True)
g1["dp_str"] = str(S.wagnerFisher('AGRGA', 'AGGGAA', True))
es1 = g1["es"][0]
g1["patch_cases"] = [[x, list(S.patching(es1, x))] for x in ['AGRGA', 'AGRGC', 'AGRGAUU', 'AGR', '', 'AGRGAA', 'UUUUU']]
G["G1"] = g1

# ---- test_input.xml ---------------------------------------------------------------------------
seqs = import_xml('test_input.xml')
ids = list(seqs.keys())
G["xml_ids"] = ids
G["xml_seqs"] = [seqs[i] for i in ids]
named = {}
for (x, y, user) in [(3, 2, False), (1, 2, False), (1, 2, True), (12, 13, False), (24, 25, True),
                     (24, 25, False), (2, 1, False), (9, 8, True)]:
    named[f"{x}->{y}:{'user' if user else 'default'}"] = full_case(
        seqs[f"piR-ocu-{x}"], seqs[f"piR-ocu-{y}"], user, want_dp=False)
G["xml_named"] = named
allp = []
for i, x in enumerate(ids):
    for j, y in enumerate(ids):
        if i == j:
            continue
        row = [i, j]
        for user in (False, True):
            dp = S.wagnerFisher(seqs[x], seqs[y], user)
            row.append(dp[-1][-1].value)
        allp.append(row)
G["xml_all_pairs"] = allp          # [i, j, dist_default, dist_user] for all 600 ordered pairs

# ---- random differential set: small pairs, everything -----------------------------------------
rnd = random.Random(20260001)
small = []
for t in range(260):
    alpha = NUC if t % 3 else NUC[:4]
    la, lb = rnd.randint(1, 11), rnd.randint(1, 11)
    a = ''.join(rnd.choices(alpha, k=la)); b = ''.join(rnd.choices(alpha, k=lb))
    if t % 5 == 0:  # homologous pair
        b = ''.join(ch if rnd.random() > 0.25 else rnd.choice(alpha) for ch in a)[:max(1, lb)]
    small.append(full_case(a, b, bool(t & 1)))
G["small"] = small

medium = []
for t in range(80):
    alpha = NUC if t % 2 else NUC[:4]
    la, lb = rnd.randint(14, 40), rnd.randint(14, 40)
    a = ''.join(rnd.choices(alpha, k=la))
    if t % 4 < 2:
        b = []
        for ch in a:
            r = rnd.random()
            if r < 0.08: b.append(rnd.choice(alpha))
            elif r < 0.12: b.extend([ch, rnd.choice(alpha)])
            elif r < 0.16: pass
            else: b.append(ch)
        b = ''.join(b) or 'A'
    else:
        b = ''.join(rnd.choices(alpha, k=lb))
    medium.append(full_case(a, b, bool((t >> 1) & 1), path_cap=3000, keep_paths=3, want_dp=(t < 20)))
G["medium"] = medium

# empty-string behaviour (SED:146-182 works; generate_es raises IndexError SED:278)
empties = []
for a, b in [('', ''), ('', 'AGU'), ('ACG', ''), ('A', '')]:
    dp = S.wagnerFisher(a, b, False)
    rec = {"a": a, "b": b, "dp_repr": [[repr(x.value) for x in row] for row in dp], "mask": tie_mask(dp)}
    paths = S.create_paths(dp)
    rec["paths_cells"] = [[(nd.i + 1, nd.j + 1) for nd in p] for p in paths]
    try:
        es = S.generate_es(paths[0], a, b)
        rec["es"] = es
    except IndexError as e:
        rec["es_error"] = "IndexError"
    empties.append(rec)
G["empties"] = empties

# symbol handling (SED:79-87)
sym = []
for a, b in [('a', 'A'), ('a', 'G'), ('T', 'T'), ('T', 'A'), ('A', 'T'), ('AT', 'AA'), ('ag', 'AG'), ('Ag', 'aG')]:
    try:
        dp = S.wagnerFisher(a, b, False)
        sym.append({"a": a, "b": b, "dp_repr": [[repr(x.value) for x in row] for row in dp]})
    except KeyError as e:
        sym.append({"a": a, "b": b, "error": "KeyError", "key": e.args[0]})
G["symbols"] = sym

# hand-edited / mismatched scripts through the sequential patch (SED:402-457)
odd = []
for t in range(120):
    c = rnd.choice(small)
    if "es" not in c:
        continue
    es = c["es"][0]
    if t % 3 == 0:
        es = S.generate_rev_es(es)
    base = S.generate_sequence_from_es(es)
    x = base + ''.join(rnd.choices(NUC, k=rnd.randint(0, 4)))
    if t % 4 == 1 and x:
        k = rnd.randrange(len(x)); x = x[:k] + rnd.choice(NUC) + x[k + 1:]
    if t % 7 == 3:
        x = x[:max(0, len(base) - 1)]
    if t % 11 == 5:   # shuffled script: arbitrary hand edit
        es = list(es); rnd.shuffle(es)
    odd.append({"es": es, "x": x, "out": list(S.patching(es, x))})
G["patch_odd"] = odd

# ---- G6: fp64 accumulation-order fingerprints ----------------------------------------------------
random.seed(0)
g6 = []
for L in (100, 200, 300, 500, 1000):
    s1 = ''.join(random.choices(NUC, k=L)); s2 = ''.join(random.choices(NUC, k=L))
    dp = S.wagnerFisher(s1, s2)
    g6.append({"L": L, "a": s1, "b": s2, "distance": dp[-1][-1].value})
    del dp
G["G6"] = g6

# ---- G7: search + top-k (IR:443-477 with a stub collection; performance.py:12-15) ---------------
class Coll:
    def __init__(self, docs): self.docs = docs
    def find(self, flt): return iter(self.docs)

coll = Coll([{"sequence": s} for s in G["xml_seqs"]])
g7 = []
for q in ['AAAAAAAAAACUCACCAUGCUGAAAAGC', G["xml_seqs"][0], 'GGGAAAUUUCCC', 'AAAAAAAAAAAGNGCUACGACAUUUGG']:
    scores = IR.search_collection(q, 'tf', coll, IR.wf_score)
    top = sorted(scores, key=itemgetter(1), reverse=True)[:6]
    g7.append({"query": q, "scores": [list(x) for x in scores], "top6": [list(x) for x in top]})
G["G7"] = g7
G["wf_score_user"] = [[a, b, IR.wf_score(a, b, True)] for a, b in [('AGRGA', 'AGGGAA'), ('ACGU', 'UGCA')]]

with open(OUT, "w") as f:
    json.dump(G, f, separators=(",", ":"))
print("wrote", OUT, os.path.getsize(OUT), "bytes")
