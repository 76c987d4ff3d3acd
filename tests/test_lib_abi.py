"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/mxe.h
declares; the product path fails loudly without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

import ntjoin_b200
from ntjoin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "mxe.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mxe_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_lib.SYMBOLS)


def test_library_exports_every_symbol():
    lib = ntjoin_b200.load_library()
    for s in header_functions():
        assert getattr(lib, s) is not None
    assert b"sm_100a" in lib.mxe_version()


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ntjoin_b200.MxeError) as ei:
        ntjoin_b200.Engine(0)
    assert "no CPU fallback" in str(ei.value)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under ntjoin_b200/ or bin/ may reference it"""
    bad = []
    for base in ("ntjoin_b200", "bin"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", "indexlr")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle_lib|libmxo|mxo_|oracle/", t):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
