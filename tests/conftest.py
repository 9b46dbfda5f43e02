import base64
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PROGRAMS = os.path.join(ROOT, "programs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_vectors():
    vecs = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))
    for v in vecs:
        v["input"] = base64.b64decode(v["input"])
        v["output"] = base64.b64decode(v["output"])
    return vecs


def program_source(name):
    return open(os.path.join(PROGRAMS, name + ".kex"), encoding="utf-8").read()


def sample(name):
    return open(os.path.join(GOLDEN, name), "rb").read()


def vec_matches(v, out):
    if out is None:
        return False
    if v["rstrip"]:
        return out.rstrip(b"\n") == v["output"].rstrip(b"\n")
    return out == v["output"]


@pytest.fixture(scope="session")
def vectors():
    return load_vectors()
