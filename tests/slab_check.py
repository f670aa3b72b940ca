"""Manual check of the slab rank path (csrc/mbd_slab.cuh) against the CPU oracle and the part pipeline:
    python tests/slab_check.py        (GPU box, repo root)
Every case runs J = 2 counts, J = 3 counts and the rank output through both paths (SD_MBD_PATH=parts | default)."""
import os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from statdepth_b200._engine import get_engine
from oracle import cpu_oracle as oracle
oracle.build()
eng = get_engine()
rng = np.random.default_rng(7)


def rows(kind, T, n):
    if kind == "normal":
        return rng.standard_normal((T, n))
    if kind == "walk":
        return np.cumsum(rng.standard_normal((T, n)), axis=0)
    if kind == "uniform":
        return rng.random((T, n))
    if kind == "expo":
        return rng.standard_exponential((T, n))
    if kind == "t3":
        return rng.standard_t(3, (T, n))
    if kind == "cauchy":
        return rng.standard_cauchy((T, n))
    if kind == "round4":
        return np.round(rng.standard_normal((T, n)), 4)
    if kind == "round2":
        return np.round(rng.standard_normal((T, n)), 2)
    if kind == "ints":
        return np.round(rng.standard_normal((T, n)) * 100)
    if kind == "const":
        return np.full((T, n), 3.25)
    if kind == "outlier":
        X = rng.standard_normal((T, n)); X[:, ::997] *= 1e9; return X
    if kind == "shifted":
        return 1e6 + 1e-3 * rng.standard_normal((T, n))
    if kind == "mixed":   # different kinds of rows in one call
        X = rng.standard_normal((T, n)); X[1] = np.round(X[1] * 10); X[2] = 7.0; return X
    if kind == "tail_all":   # every row fails the hist kernel's validation: masked part pipeline for all rows
        return rng.standard_cauchy((T, n))
    if kind == "tail_many":  # 10 of 40 rows: slab ranks 30 rows, the part pipeline the other 10
        X = rng.standard_normal((T, n)); X[5:15] = rng.standard_cauchy((10, n)); return X
    if kind == "tail_few":   # 2 of 40 rows: generic path for those
        X = rng.standard_normal((T, n)); X[7] = rng.standard_cauchy(n); X[30] = rng.standard_cauchy(n); return X
    raise ValueError(kind)


bad = cases = 0
t0 = time.time()
for n in (16384, 20000, 50000, 100000, 131072):
    for kind in ("normal", "walk", "uniform", "expo", "t3", "cauchy", "round4", "round2", "ints", "const", "outlier",
                 "shifted", "mixed", "tail_all", "tail_many", "tail_few"):
        T = 4 if n <= 50000 else 3
        if kind == "mixed":  # 2 unfit rows of 40: below the 1/16 threshold, so the slab path keeps the block
            if n > 20000:
                continue
            T = 40
        if kind.startswith("tail_"):
            if n not in (20000, 100000):
                continue
            T = 40
        X = rows(kind, T, n)
        want2, wb, wa = oracle.mbd_counts_all(X, j=2, want_ranks=True)
        want3 = oracle.mbd_counts_all(X, j=3)
        res = {}
        for path in ("parts", "slab"):
            if path == "parts":
                os.environ["SD_MBD_PATH"] = "parts"
            else:
                os.environ.pop("SD_MBD_PATH", None)
            got2 = eng.band_depth_counts(X, None, 2, True)
            tm = eng.timings()
            got3 = eng.band_depth_counts(X, None, 3, True)
            b, a = eng.band_ranks(X)
            ok = (got2 == want2).all() and (got3 == want3).all() and (b == wb).all() and (a == wa).all()
            res[path] = (ok, tm["launches"], tm["fallback_rows"], tm["kernel_ns"] / 1e3)
            cases += 1
            if not ok:
                bad += 1
                print("MISMATCH", path, kind, n, "j2 diff", int((got2 != want2).sum()), "j3 diff", int((got3 != want3).sum()),
                      "ranks diff", int((b != wb).sum()), int((a != wa).sum()), flush=True)
        print("%-8s n=%6d  parts: ok=%s launches=%d fb=%d %.0fus | slab: ok=%s launches=%d fb=%d %.0fus" %
              ((kind, n) + res["parts"] + res["slab"]), flush=True)
print("cases", cases, "bad", bad, "seconds %.1f" % (time.time() - t0))
sys.exit(1 if bad else 0)
