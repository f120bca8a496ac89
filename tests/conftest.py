import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from checkers import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from checkers import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libtrico_ref.so not built (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def golden():
    import gzip
    import json
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with gzip.open(os.path.join(here, "kat.json.gz"), "rt") as f:
        kat = json.load(f)
    bunny = dict(np.load(os.path.join(here, "bunny_head.npz")))
    with open(os.path.join(here, "bunny_facts.json")) as f:
        facts = json.load(f)
    full = dict(np.load(os.path.join(here, "bunny_full.npz")))
    return {"kat": kat, "bunny": bunny, "facts": facts, "bunny_full": full}
