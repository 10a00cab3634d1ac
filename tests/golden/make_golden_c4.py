#!/usr/bin/env python
"""Generate tests/golden/c4_digest.json: the canonical edit script (= create_paths(dp)[0], SED:228-271, rule
SURVEY a8) of the BASELINE config 4 pair (50 kb x 50 kb, seed 20260004) under both shipped cost tables, as
computed by the CPU oracle (oracle/wf_oracle.c, itself pinned to the unmodified reference by ref_golden.json).
Recorded: distance, number of ops and SHA-256 of the op / oi / oj arrays — enough to check the GPU long-pair
path, whose 32-bit keys wrap at this size, bit for bit on the GPU box where the oracle run would take too long
for a test.  Also a 30 kb x 50 kb independent random pair (a different key regime: cost grows with the path).

    python tests/golden/make_golden_c4.py        (about a minute, ~1.3 GB of RAM)"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from _synth import c4_pair  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c4_digest.json")


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def record(a, b, costs):
    ops, oi, oj, d = O.canonical_script(O.decode(a), O.decode(b), costs)
    return {"m": int(a.shape[0]), "n": int(b.shape[0]), "dist": float(d).hex(), "n_ops": int(ops.shape[0]),
            "op_sha256": digest(ops.astype(np.uint8)), "oi_sha256": digest(oi.astype(np.int32)),
            "oj_sha256": digest(oj.astype(np.int32))}


def main():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_golden.json")))
    out = {}
    a, b = c4_pair()
    out["c4_default_costs"] = record(a, b, g["default_costs"])
    out["c4_user_costs"] = record(a, b, g["user_costs"])
    rng = np.random.default_rng(20260044)
    a2 = rng.integers(0, 4, size=30000, dtype=np.uint8); b2 = rng.integers(0, 4, size=50000, dtype=np.uint8)
    out["random_30k_50k_user_costs"] = record(a2, b2, g["user_costs"])
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
