"""If a REAL btllib `indexlr` is reachable (on PATH or under baseline/_ref/), byte-compare its TSVs with the oracle's
and with the engine's: the only way to pin what the reference's own fixtures leave open (SURVEY.md section 0.5 --
out_hash values under the `sum` combiner, tie direction, N handling, lower case).  Skipped when there is none (this
container and the GPU boxes have neither btllib nor network); the probe itself always runs and is recorded."""
import glob
import os
import shutil
import subprocess

import pytest

import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ["ref.fa", "ref.multiple.fa", "scaf.f-f.fa", "scaf.f-f.termN.unassigned.fa", "scaf.more_seqs.fa"]


def find_real_indexlr():
    """a btllib indexlr that is not this repo's drop-in (bin/indexlr) nor the oracle's CLI"""
    ours = {os.path.realpath(os.path.join(ROOT, "bin", "indexlr")), os.path.realpath(oracle_lib.CLI)}
    cands = [shutil.which("indexlr")] + glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "indexlr"), recursive=True)
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK) and os.path.realpath(c) not in ours:
            try:
                head = open(c, "rb").read(4096)
            except OSError:
                continue
            if b"ntjoin_b200" in head or b"mxo_indexlr" in head:        # a wrapper of ours
                continue
            return c
    return None


def test_probe_runs():
    exe = find_real_indexlr()
    assert exe is None or os.path.isfile(exe)


@pytest.mark.skipif(find_real_indexlr() is None, reason="no real btllib indexlr on PATH or under baseline/_ref/")
@pytest.mark.parametrize("fname", FIXTURES)
@pytest.mark.parametrize("kw", [(32, 1000), (32, 500), (15, 10)])
def test_oracle_vs_real_indexlr(fname, kw, tmp_path):
    k, w = kw
    fa = os.path.join(ROOT, "tests", "golden", "inputs", fname)
    real = subprocess.check_output([find_real_indexlr(), "--seq", "--long", "--pos", "-k", str(k), "-w", str(w), "-t", "4", fa])
    mine = subprocess.check_output([oracle_lib.CLI, "--seq", "--long", "--pos", "-k", str(k), "-w", str(w), fa])
    assert mine == real, "oracle/mxo.c differs from the real indexlr: one of the unpinned semantics (SURVEY 0.5) is wrong"


@pytest.mark.gpu
@pytest.mark.skipif(find_real_indexlr() is None, reason="no real btllib indexlr on PATH or under baseline/_ref/")
@pytest.mark.parametrize("fname", FIXTURES)
def test_engine_vs_real_indexlr(engine, fname, tmp_path):
    fa = os.path.join(ROOT, "tests", "golden", "inputs", fname)
    real = subprocess.check_output([find_real_indexlr(), "--seq", "--long", "--pos", "-k", "32", "-w", "500", "-t", "4", fa])
    sk = engine.sketch_file(fa, 32, 500)
    out = tmp_path / "o.tsv"
    sk.write_tsv(out, pos=True, strand=False, seq=True)
    assert out.read_bytes() == real
