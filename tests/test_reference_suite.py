"""CPU: the reference's OWN test-suite (tests/ntjoin_test.py, 20 tests) run through its unmodified make driver and
bin/*.py, with the CPU oracle standing in for `indexlr` / btllib and small stand-ins for the third-party packages that
are not installed here (tests/harness/).  Pins the oracle's `sum` canonical hash, window rule and
steps 2-3 on every scaffold path, gap size, orientation, AGP line and overlap cut point the reference asserts
(SURVEY.md 8(c) item 6).  Needs the reference checkout: skipped where /root/reference does not exist (the GPU box)."""
import os
import shutil
import subprocess
import sys

import pytest

import oracle_lib

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


needs_reference = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "tests", "ntjoin_test.py")),
                                     reason="reference checkout not on this box")


def run_suite(tmp_path, canonical=None, extra=(), dropin=False):
    oracle_lib.build()
    work = tmp_path / "ntJoin"
    shutil.copytree(REF, work, symlinks=True, ignore=shutil.ignore_patterns(".git"))
    env = dict(os.environ)
    env["PATH"] = os.path.join(HERE, "harness", "bin") + os.pathsep + env["PATH"]
    env["PYTHONPATH"] = os.path.join(HERE, "harness", "modules") + os.pathsep + HERE + os.pathsep + env.get("PYTHONPATH", "")
    env["MXO_INDEXLR"] = oracle_lib.CLI
    env.pop("NTJOIN_B200", None)
    env.pop("MXO_CANONICAL", None)
    if canonical:
        env["MXO_CANONICAL"] = canonical
    if dropin:      # the product's drop-in layer on, its two engine calls served by the oracle (tests/harness/fake_engine.py)
        env["NTJOIN_B200"] = "1"
        env["MXE_REPO_ROOT"] = os.path.dirname(HERE)
        env["PYTHONPATH"] = os.path.join(HERE, "harness", "dropin_site") + os.pathsep + env["PYTHONPATH"]
        env["MXE_DROPIN_TRACE"] = str(tmp_path / "dropin_trace.txt")
    r = subprocess.run([sys.executable, "-m", "pytest", "ntjoin_test.py", "-q", "-p", "no:cacheprovider", *extra],
                       cwd=work / "tests", env=env, capture_output=True, text=True)
    return r, (r.stdout + r.stderr)[-3000:]


@needs_reference
@pytest.mark.timeout(1500)
def test_reference_suite_with_oracle_sketch(tmp_path):
    r, tail = run_suite(tmp_path, extra=("-x",))
    assert r.returncode == 0, tail
    assert "20 passed" in r.stdout, tail


@needs_reference
@pytest.mark.timeout(1500)
def test_reference_suite_rejects_legacy_min_combiner(tmp_path):
    """the suite is discriminating: with the legacy canonical hash min(fwd, rev) -- the one the shipped golden TSVs
    were made with -- several of the reference's tests fail, so it is `fwd + rev` that current ntJoin expects"""
    r, tail = run_suite(tmp_path, canonical="min")
    assert r.returncode != 0 and " failed" in r.stdout, tail
    assert "test_regions_ff_rr" in r.stdout and "overlap" in r.stdout, tail


@needs_reference
@pytest.mark.timeout(1500)
def test_reference_suite_with_dropin_layer(tmp_path):
    """the reference's 20 tests with the PRODUCT's seam-S3 layer switched on through its own import hook:
    read_minimizers / filter_minimizers / build_graph rebuilt from arrays, Ntjoin.print_graph through mxe_write_dot,
    NtjoinScaffolder.find_mx_min_max from arrays, plain-list callers (ntjoin_overlap) falling through -- only the two
    engine calls underneath are served by the oracle here, because this container has no GPU"""
    r, tail = run_suite(tmp_path, extra=("-x",), dropin=True)
    assert r.returncode == 0, tail
    assert "20 passed" in r.stdout, tail
    # the hooks really ran: every seam function was served through the drop-in layer in every run, and the plain-list
    # callers of the overlap code (bin/ntjoin_overlap.py:25-28,132) fell through to the reference's originals
    trace = (tmp_path / "dropin_trace.txt").read_text().split("\n")
    for fn in ("read_minimizers", "filter_minimizers", "build_graph", "print_graph", "find_mx_min_max"):
        assert trace.count(fn + " engine") >= 20, (fn, trace.count(fn + " engine"))
    assert trace.count("build_graph original") > 0 and trace.count("filter_minimizers original") > 0
    # steps 2-3 run once per set of assemblies: build_graph reuses the pass filter_minimizers made
    assert trace.count("filter_and_edges cached") >= 20 and trace.count("filter_and_edges engine") >= 20
    assert trace.count("print_graph original") == 0 and trace.count("find_mx_min_max original") == 0
    dots = sorted((tmp_path / "ntJoin" / "tests").glob("*.mx.dot"))
    assert dots
    for d in dots:      # written by the array writer from the engine-built graph: vertices in ascending hash order
        names = [int(line.split('"')[1]) for line in d.read_text().splitlines() if "[label=" in line]
        assert names == sorted(names), d
