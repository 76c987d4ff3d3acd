import time, numpy as np, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import ntjoin_b200
from ntjoin_b200 import synth
eng = ntjoin_b200.Engine(0)
seq, offs, _ = synth.make_reference(300000, n_chrom=3, seed=5)
sk = eng.sketch_buffers(seq, offs, 32, 100)
for variant in (0, 1, 1):
    eng.set_option("filter_variant", variant)
    ts = []
    for i in range(6):
        t0 = time.perf_counter()
        res = eng.filter_and_edges([sk, sk], [1.0, 1.0])
        res.counts()
        ts.append(time.perf_counter() - t0)
        res.close()
    print("variant", variant, ["%.4f" % t for t in ts])
eng.close()
