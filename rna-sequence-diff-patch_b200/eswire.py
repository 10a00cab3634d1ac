"""eswire.py — edit-script wire formats.

* the reference GUI's JSON export/import: {'edit_script': [op dicts]} with indent=4 (gui.py:629-639,
  655-664) — byte-compatible so files travel between the reference GUI and this package;
* packed scripts (rsd_script_batch: op codes + entered cells) <-> the reference's op dicts (SED:284-286);
* generate_rev_es (SED:338-369) directly on packed scripts."""
from __future__ import annotations

import json

import numpy as np

from . import sed

_OPCODE = {"insert": 0, "delete": 1, "update": 2}


def dump_es(es, path: str):
    """gui.py:636-639."""
    with open(path, "w") as f:
        json.dump({"edit_script": es}, f, indent=4)


def load_es(path: str):
    """gui.py:655-664."""
    with open(path) as f:
        return json.load(f)["edit_script"]


def packed_to_es(op, oi, oj, str1: str, str2: str):
    return sed.es_from_packed(op, oi, oj, str1, str2)


def es_to_packed(es):
    """op dicts (of a generated script) -> (op uint8, oi int32, oj int32): entered cell = index + 1."""
    op = np.array([_OPCODE[e["operation"]] for e in es], dtype=np.uint8)
    oi = np.array([e["source"]["index"] + 1 for e in es], dtype=np.int32)
    oj = np.array([e["destination"]["index"] + 1 for e in es], dtype=np.int32)
    return op, oi, oj


def rev_packed(op, oi, oj):
    """The packed script that turns B back into A: insert <-> delete, entered cells transposed.
    Valid input for Engine.patch_batch(…, A=B, B=A, X=…)."""
    op = np.asarray(op)
    rop = np.where(op == 0, 1, np.where(op == 1, 0, 2)).astype(np.uint8)
    return rop, np.asarray(oj, dtype=np.int32).copy(), np.asarray(oi, dtype=np.int32).copy()


def rev_es_from_packed(op, oi, oj, str1: str, str2: str):
    """generate_rev_es(es_from_packed(op, oi, oj, str1, str2)) built directly (SED:338-369), including
    the reference's quirk that a reversed delete becomes an insert whose source index is
    destination.index - 1 while its character stays the destination character (SED:356)."""
    out = []
    for o, i, j in zip(op, oi, oj):
        i, j = int(i), int(j)
        src = {"character": str2[j - 1], "index": j - 1}
        dst = {"character": str1[i - 1], "index": i - 1}
        if o == 0:
            out.append({"operation": "delete", "source": src, "destination": dst})
        elif o == 1:
            out.append({"operation": "insert", "source": {"index": j - 2, "character": str2[j - 1]}, "destination": dst})
        else:
            out.append({"operation": "update", "source": src, "destination": dst})
    return out
