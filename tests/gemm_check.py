"""Standalone check of the tcgen05 Gram kernel (run under `timeout`: a bad descriptor can hang)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu_oracle
from statdepth_b200 import _engine as E

eng = E.Engine(0)
eng.set_option(E.OPT_BD_IMPL, E.BD_GEMM)
ok = True
for (T, n, seed, kind) in [(64, 128, 0, "walk"), (64, 256, 1, "walk"), (100, 200, 2, "walk"), (5, 6, 3, "walk"),
                           (130, 300, 4, "round"), (512, 1000, 5, "walk"), (70, 600, 6, "noncross")]:
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, n)).cumsum(0)
    if kind == "round":
        X = np.round(X)
    if kind == "noncross":
        X = np.outer(rng.random(T) + 0.1, rng.random(n))
    t0 = time.time()
    got = eng.band_depth_counts(X, None, 2, False)
    exp = cpu_oracle.bd_counts(X)
    good = bool((got == exp).all())
    ok &= good
    print(T, n, kind, "OK" if good else "MISMATCH", "first", got[:4], exp[:4], "%.3fs" % (time.time() - t0), flush=True)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
