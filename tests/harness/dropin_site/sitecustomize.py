"""sitecustomize for the CPU run of the reference's test-suite WITH the product's drop-in layer (tests only).

Python imports one `sitecustomize`; this one (1) executes the product's dropin/sitecustomize.py, which registers the
import hook when NTJOIN_B200=1, and (2) presets the drop-in layer's engine with the oracle-backed stand-in, because this
container has no GPU.  The product has no switch for this: the preset happens from outside, here.
"""
import os
import runpy
import sys

_REPO = os.environ.get("MXE_REPO_ROOT")
if _REPO and os.environ.get("NTJOIN_B200", "0") not in ("", "0"):
    for p in (_REPO, os.path.join(_REPO, "tests"), os.path.join(_REPO, "tests", "harness")):
        if p not in sys.path:
            sys.path.insert(0, p)
    runpy.run_path(os.path.join(_REPO, "dropin", "sitecustomize.py"))
    import ntjoin_b200.dropin as _dropin
    import fake_engine as _fake
    _dropin._ENGINE = _fake.FakeEngine()
