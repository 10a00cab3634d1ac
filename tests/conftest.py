import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "ref_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def default_costs(golden):
    return golden["default_costs"]


@pytest.fixture(scope="session")
def user_costs(golden):
    return golden["user_costs"]


@pytest.fixture(scope="session")
def dropin(golden, tmp_path_factory):
    """The drop-in modules imported the way the reference's are: from a working directory that holds costs.json
    and user_costs.json (SED:6-18 reads them from the CWD at import), drop-in directory first on sys.path."""
    drop = os.path.join(ROOT, "rna-sequence-diff-patch_b200", "dropin")
    work = tmp_path_factory.mktemp("refcwd")
    with open(work / "costs.json", "w") as f:
        json.dump(golden["default_costs"], f)
    with open(work / "user_costs.json", "w") as f:
        json.dump(golden["user_costs"], f)
    cwd = os.getcwd()
    os.chdir(work)
    sys.path.insert(0, drop)
    try:
        for name in ("StringEditDistance", "IRMethods"):
            sys.modules.pop(name, None)
        import StringEditDistance
        import IRMethods
    finally:
        os.chdir(cwd)

    class Mods:
        SED, IR, cwd = StringEditDistance, IRMethods, str(work)
    return Mods
