#!/usr/bin/env python
"""Generate tests/golden/ref_golden_ir.json by running the UNMODIFIED reference's set / multiset /
TF-vector similarity search (IRMethods.py IR:49-389 through search_collection IR:443-477).

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden_ir.py

The collection is a stub exposing find({}) -> documents {'sequence', 'tf'} where 'tf' is the pickled
convert_to_tf_vector(sequence), exactly what the reference's importer stores (fa_import.py:49).
Scores are recorded with float.hex() so they round-trip bit for bit ('nan' for the reference's 0/0).
Nothing from the reference is copied into the repository except these input/output vectors."""
import contextlib
import io
import json
import math
import os
import pickle
import random
import sys
import warnings

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_ir.json")
NUC = ['A', 'G', 'C', 'U', 'Y', 'R', 'W', 'S', 'K', 'M', 'D', 'V', 'H', 'B', 'N']  # IR:13

sys.dont_write_bytecode = True
os.chdir(REF)
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    import IRMethods as IR
    from import_xml import import_xml

warnings.simplefilter("ignore")


class Coll:
    def __init__(self, seqs):
        self.docs = [{'sequence': s, 'tf': pickle.dumps(IR.convert_to_tf_vector(s))} for s in seqs]

    def find(self, _):
        return iter(self.docs)


def hx(v):
    v = float(v)
    return "nan" if math.isnan(v) else v.hex()


METHODS = ["set_intersection_similarity", "set_jaccard_similarity", "set_dice_similarity",
           "multi_intersection_similarity", "multi_jaccard_similarity", "multi_dice_similarity",
           "cosine", "pearson", "euclidian_distance", "manhattan_distance", "tanimoto_distance", "dice_dist"]


def main():
    rnd = random.Random(20260418)
    xml = list(import_xml(os.path.join(REF, "test_input.xml")).values())       # file order (dict insertion order)
    docs = list(xml)
    for L in [1, 2, 2, 3, 5, 8, 13, 24, 31, 31, 40, 64, 100, 300]:
        docs.append(''.join(rnd.choices(NUC, k=L)))
    for L in [2, 7, 24, 28, 31, 150]:
        docs.append(''.join(rnd.choices(NUC[:4], k=L)))
    for _ in range(20):                                              # mostly ACGU with a few ambiguity codes
        s = rnd.choices(NUC[:4], k=rnd.randint(20, 36))
        for _ in range(rnd.randint(0, 3)):
            s[rnd.randrange(len(s))] = rnd.choice(NUC[4:])
        docs.append(''.join(s))
    queries = [xml[8], 'AAAAAAAAAACUCACCAUGCUGAAAAGC', docs[len(xml) + 8], docs[len(xml) + 12], 'AN', 'ACGU', 'G',
               ''.join(rnd.choices(NUC, k=29))]
    coll = Coll(docs)
    out = {"docs": docs, "queries": queries, "methods": METHODS, "scores": {}, "tf": {}, "multiset": {}}
    for q in queries:
        out["scores"][q] = {}
        for name in METHODS:
            res = IR.search_collection(q, 'tf', coll, getattr(IR, name))
            assert [s for s, _ in res] == docs
            out["scores"][q][name] = [hx(v) for _, v in res]
    for s in docs[len(xml):len(xml) + 14] + queries:                 # representation vectors themselves
        out["tf"][s] = [hx(v) for v in IR.convert_to_tf_vector(s).reshape(-1)]
        out["multiset"][s] = [hx(v) for v in IR.convert_to_multi_set(s)]
    # ---- idf / tf-idf vectors (IR:189-223) and the search over them (IR:458-465): a collection whose documents
    # carry the 'idf' vector update_idfs() would have stored (fa_import.py:30-36, commented out in the reference)
    small = docs[len(xml) + 6:len(xml) + 14] + docs[-6:] + xml[:6]
    class CollIdf:
        def __init__(self, seqs):
            self.docs = [{'sequence': s} for s in seqs]
            for d in self.docs:
                d['tf'] = pickle.dumps(IR.convert_to_tf_vector(d['sequence']))
            for d in self.docs:
                d['idf'] = pickle.dumps(IR.convert_to_idf_vector(d['sequence'], self))
        def find(self, _):
            return iter(self.docs)
        def count_documents(self, _):
            return len(self.docs)
    cidf = CollIdf(small)
    out["idf_docs"] = small
    out["idf"] = {d['sequence']: [hx(v) for v in pickle.loads(d['idf']).reshape(-1)] for d in cidf.docs}
    out["idf_pairlist"] = {}
    for a, b in [(small[0], small[1]), (small[3], small[9]), ('AN', 'ACGU')]:       # gui.py:480-484 shape
        out["idf_pairlist"][a + "|" + b] = {
            "idf_a": [hx(v) for v in IR.convert_to_idf_vector(a, list_of_docs=[a, b]).reshape(-1)],
            "tfidf_a": [hx(v) for v in IR.create_tf_idf_vector(a, list_of_docs=[a, b]).reshape(-1)],
            "tfidf_b": [hx(v) for v in IR.create_tf_idf_vector(b, list_of_docs=[a, b]).reshape(-1)]}
    out["idf_scores"] = {}
    for q in [small[2], queries[1], 'AN']:
        out["idf_scores"][q] = {}
        for vt in ('idf', 'tf-idf'):
            out["idf_scores"][q][vt] = {}
            for name in METHODS[6:]:
                res = IR.search_collection(q, vt, cidf, getattr(IR, name))
                out["idf_scores"][q][vt][name] = [hx(v) for _, v in res]
    out["compare_pair"] = [[p, s, hx(IR.compare_pair_to_seq(p, s, is_document=False))]
                           for p in ('AA', 'AN', 'NN', 'YR', 'GU', 'BD') for s in small[:8]]
    out["possibilities"] = {c: [IR.possibilities(c)[0], [hx(v) for v in IR.possibilities(c)[1]]] for c in NUC}
    with open(OUT, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(docs), "docs,", len(queries), "queries")


if __name__ == "__main__":
    main()
