"""Drop-in for the reference's StringEditDistance.py (put this directory first on sys.path).

Same names, argument meaning and error behaviour as /root/reference/StringEditDistance.py; the
matrix fill runs on the GPU through librsd.so.  Like the reference (SED:6-18) the cost files are
read from the current working directory at import: `costs.json` (when it is absent the identical table
built into cost_tables.py is used, where the reference would raise FileNotFoundError), and
`user_costs.json` — when THAT is absent user_costs is the default table itself (the same object) and the
reference's notice is printed (SED:12-18).  The reference's import-time self-test print (SED:463-471) is
not reproduced: call selftest()."""
import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from rna_sequence_diff_patch_b200 import sed as _sed  # noqa: E402
from rna_sequence_diff_patch_b200.sed import Edge, Node  # noqa: E402,F401


from rna_sequence_diff_patch_b200 import cost_tables as _tables  # noqa: E402


if os.path.exists("costs.json"):
    with open("costs.json", "r") as f:
        default_costs = json.load(f)
else:
    default_costs = _tables.default_costs()

try:
    with open("user_costs.json", "r") as f:
        user_costs = json.load(f)
except (OSError, IOError):
    user_costs = default_costs                       # SED:16: an alias of the default table, not a copy
    print('Could not find user costs file')


def reload_user_costs():
    """SED:24-27."""
    global user_costs
    with open('user_costs.json', 'r') as f:
        user_costs = json.load(f)


def _table(userCosts):
    return user_costs if userCosts else default_costs


def cost(char1, char2, userCosts=False):
    return _sed.cost(char1, char2, _table(userCosts))


def min_cost(dp, i, j, str1, str2, userCosts=False):
    return _sed.min_cost(dp, i, j, str1, str2, _table(userCosts))


def wagnerFisher(str1, str2, userCosts=False):
    return _sed.wagner_fisher(str1, str2, _table(userCosts))


def create_paths(dp):
    return _sed.create_paths(dp)


def generate_es(path, str1, str2):
    return _sed.generate_es(path, str1, str2)


def generate_rev_es(es):
    return _sed.generate_rev_es(es)


def generate_sequence_from_es(es):
    return _sed.generate_sequence_from_es(es)


def patching(es, str1):
    return _sed.patching(es, str1)


def selftest():
    """What the reference prints at import (SED:463-471)."""
    str1, str2 = 'AGRGA', 'AGGGAA'
    dp = wagnerFisher(str1, str2, True)
    print(dp)
    for path in create_paths(dp):
        print(generate_es(path, str1, str2))
