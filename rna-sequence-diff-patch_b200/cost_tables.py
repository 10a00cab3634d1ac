"""cost_tables.py — the reference's shipped cost tables as data (costs.json:1-260, user_costs.json:1).

The reference keeps its substitution costs in two JSON files read from the current directory
(StringEditDistance.py:6-18).  The drop-in still reads those files when they are present; this module is
what it falls back to, and what the tests, tools and bench.py use, so the package needs no data files.
DEFAULT_COSTS / USER_COSTS have exactly the JSON layout {"insert": f, "delete": f, "update": {src: {dst: f}}}
with rows = source symbol and columns = destination symbol in the order of IRMethods.py:13; the values are
pinned against the reference's own files by tests/golden/ref_golden.json (test_oracle_golden.py)."""
from __future__ import annotations

import copy
import json

SYMBOLS = "AGCUYRWSKMDVHBN"

# update cost, row = source symbol, column = destination symbol (asymmetric: e.g. Y->D 0.66 but D->Y 0.83)
#     A     G     C     U     Y     R     W     S     K     M     D     V     H     B     N
_DEFAULT_UPDATE = (
    (0.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 1.0, 1.0, 0.5, 0.66, 0.66, 0.66, 1.0, 0.75),   # A
    (1.0, 0.0, 1.0, 1.0, 1.0, 0.5, 1.0, 0.5, 0.5, 1.0, 0.66, 0.66, 1.0, 0.66, 0.75),   # G
    (1.0, 1.0, 0.0, 1.0, 0.5, 1.0, 1.0, 0.5, 1.0, 0.5, 1.0, 0.66, 0.66, 0.66, 0.75),   # C
    (1.0, 1.0, 1.0, 0.0, 0.5, 1.0, 0.5, 1.0, 0.5, 1.0, 0.66, 1.0, 0.66, 0.66, 0.75),   # U
    (1.0, 1.0, 0.5, 0.5, 0.0, 1.0, 0.75, 0.75, 0.75, 0.75, 0.66, 0.66, 0.66, 0.66, 0.75),   # Y
    (0.5, 0.5, 1.0, 1.0, 1.0, 0.0, 0.75, 0.75, 0.75, 0.75, 0.66, 0.66, 0.83, 0.83, 0.75),   # R
    (0.5, 1.0, 1.0, 0.5, 0.75, 0.75, 0.0, 1.0, 0.75, 0.75, 0.66, 0.83, 0.66, 0.83, 0.75),   # W
    (1.0, 0.5, 0.5, 1.0, 0.75, 0.75, 1.0, 0.0, 0.75, 0.75, 0.83, 0.66, 0.83, 0.66, 0.75),   # S
    (1.0, 0.5, 1.0, 0.5, 0.75, 0.75, 0.75, 0.75, 0.0, 1.0, 0.66, 0.83, 0.83, 0.66, 0.75),   # K
    (0.5, 1.0, 0.5, 1.0, 0.75, 0.75, 0.75, 0.75, 1.0, 0.0, 0.83, 0.66, 0.66, 0.83, 0.75),   # M
    (0.66, 0.66, 1.0, 0.66, 0.83, 0.66, 0.66, 0.83, 0.66, 0.83, 0.0, 0.78, 0.78, 0.78, 0.75),   # D
    (0.66, 0.66, 0.66, 1.0, 0.83, 0.66, 0.83, 0.66, 0.83, 0.66, 0.78, 0.0, 0.78, 0.78, 0.75),   # V
    (0.66, 1.0, 0.66, 0.66, 0.66, 0.83, 0.66, 0.83, 0.83, 0.66, 0.78, 0.78, 0.0, 0.78, 0.75),   # H
    (1.0, 0.66, 0.66, 0.66, 0.66, 0.83, 0.83, 0.66, 0.66, 0.83, 0.78, 0.78, 0.78, 0.0, 0.75),   # B
    (0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.75, 0.0),   # N
)
# the shipped user table = the default one with these overrides (user_costs.json:1)
_USER_OVERRIDES = {("A", "C"): 2.0, ("A", "K"): 1.5, ("G", "A"): 2.0, ("G", "C"): 2.0}


def _table(insert: float, delete: float, overrides=None) -> dict:
    upd = {a: {b: _DEFAULT_UPDATE[i][j] for j, b in enumerate(SYMBOLS)} for i, a in enumerate(SYMBOLS)}
    for (a, b), v in (overrides or {}).items():
        upd[a][b] = v
    return {"insert": insert, "delete": delete, "update": upd}


DEFAULT_COSTS = _table(1.0, 1.0)
USER_COSTS = _table(2.0, 3.0, _USER_OVERRIDES)


def default_costs() -> dict:
    return copy.deepcopy(DEFAULT_COSTS)


def user_costs() -> dict:
    return copy.deepcopy(USER_COSTS)


def write_cost_files(directory: str = ".") -> None:
    """Write costs.json / user_costs.json in the reference's format (for a checkout that lacks them)."""
    import os
    with open(os.path.join(directory, "costs.json"), "w") as f:
        json.dump(DEFAULT_COSTS, f, indent=2)
    with open(os.path.join(directory, "user_costs.json"), "w") as f:
        json.dump(USER_COSTS, f)
