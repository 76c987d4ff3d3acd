"""pymannkendall for the reference harness: importable; original_test is only called under --mkt (bin/ntjoin_assemble.py:38)."""
from collections import namedtuple

_Result = namedtuple("Mann_Kendall_Test", ["trend", "h", "p", "z", "Tau", "s", "var_s", "slope", "intercept"])


def original_test(x, alpha=0.05):
    x = list(x)
    n = len(x)
    s = sum((x[j] > x[i]) - (x[j] < x[i]) for i in range(n - 1) for j in range(i + 1, n))
    trend = "increasing" if s > 0 else "decreasing" if s < 0 else "no trend"
    return _Result(trend, s != 0, 0.0 if s else 1.0, float(s), 0.0, float(s), 0.0, 0.0, 0.0)
