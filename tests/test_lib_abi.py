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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """the boundary is a C ABI: include/mxe.h compiles as C99 (no C++, no torch types) and a C program links against
    libmxe.so, reads the version string and -- on a box without a GPU -- gets a clean error from mxe_create"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "use_mxe.c"
    src.write_text('''
#include <stdio.h>
#include <string.h>
#include "mxe.h"
int main(void)
{
    mxe_t* e = NULL;
    printf("%s\\n", mxe_version());
    int rc = mxe_create(0, &e);
    if (rc != MXE_OK) { printf("create: %d %s\\n", rc, mxe_last_error()); return e == NULL ? 0 : 2; }
    mxe_destroy(e);
    printf("create: ok\\n");
    return 0;
}
''')
    exe = tmp_path / "use_mxe"
    libdir = os.path.dirname(ntjoin_b200.library_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libmxe.so", "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.splitlines()
    assert "sm_100a" in lines[0]
    assert lines[1] == "create: ok" or ("create:" in lines[1] and "no CPU fallback" in lines[1])
