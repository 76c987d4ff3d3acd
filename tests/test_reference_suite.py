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


def run_suite(tmp_path, canonical=None, extra=()):
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
