"""The C-ABI library loads and exports every symbol include/kexcuda.h
declares (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from kleenexlang_b200 import build, runtime


def test_exports_match_header():
    lib_path = build.build()
    L = ctypes.CDLL(lib_path)
    hdr = open(os.path.join(ROOT, "include", "kexcuda.h")).read()
    declared = set(re.findall(r"\b(kex_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(runtime.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_in_product_path():
    # the product package must never import the oracle
    pkg = os.path.join(ROOT, "kleenexlang_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dp, f), encoding="utf-8").read()
                # the top-level `oracle` package (test infrastructure), as a module -- not the package's own
                # frontend/oracle_action.py, which restates the reference's *oracle machine* (OracleMachine.hs)
                assert not re.search(r"^\s*(import|from)\s+oracle(\.|\s|$)", txt, re.M), f
                assert "kex_oracle" not in txt and "libkexoracle" not in txt, f


def test_error_strings():
    L = runtime.lib()
    assert L.kex_strerror(0) == b"ok"
    assert b"blob" in L.kex_strerror(-1)
