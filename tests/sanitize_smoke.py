"""Small pass over every kernel for `compute-sanitizer --tool memcheck python tests/sanitize_smoke.py`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import cpu_oracle as co
from statdepth_b200 import _engine as E

eng = E.Engine(0)
rng = np.random.default_rng(0)
ok = True


def check(name, cond):
    global ok
    ok &= bool(cond)
    print(name, "OK" if cond else "MISMATCH", flush=True)


for T, n in ((7, 300), (9, 1500), (5, 3000)):
    X = rng.standard_normal((T, n)).cumsum(0)
    check("mbd %dx%d" % (T, n), (eng.band_depth_counts(X, None, 2, True) == co.mbd_counts_all(X)).all())
Xr = np.round(rng.standard_normal((6, 2500)).cumsum(0))
check("mbd ties/generic", (eng.band_depth_counts(Xr, None, 3, True) == co.mbd_counts_all(Xr, j=3)).all())
b, a = eng.band_ranks(Xr)
_, rb, ra = co.mbd_counts_all(Xr, want_ranks=True)
check("ranks", (b == rb).all() and (a == ra).all())
X = rng.standard_normal((70, 333)).cumsum(0)
check("bd bits", (eng.band_depth_counts(X, None, 2, False) == co.bd_counts(X)).all())
check("bd bits F-order", (eng.band_depth_counts(np.asfortranarray(X), [3, 1], 2, False) == co.bd_counts(X, [3, 1])).all())
eng.set_option(E.OPT_BD_IMPL, E.BD_GEMM)
check("bd gram", (eng.band_depth_counts(X, None, 2, False) == co.bd_counts(X)).all())
eng.set_option(E.OPT_BD_IMPL, E.BD_AUTO)
Xs = rng.standard_normal((20, 40)).cumsum(0)
check("bd j3", (eng.band_depth_counts(Xs, None, 3, False) == co.bd_counts(Xs, j=3)).all())
mem = (rng.random((3, 333)) < 0.5).astype(np.uint8)
qs = np.stack([rng.choice(np.flatnonzero(mem[i]), 2, replace=False) for i in range(3)])
got = eng.band_depth_counts_batched(X, mem, qs, 2, True)
exp = np.stack([co.mbd_counts_all(np.ascontiguousarray(X[:, np.flatnonzero(mem[i])]))[
    [int(np.searchsorted(np.flatnonzero(mem[i]), g)) for g in qs[i]]] for i in range(3)])
check("batched", (got == exp).all())
P = rng.standard_normal((90, 2))
check("l1", np.allclose(eng.l1_depth(P), co.l1_depth(P), rtol=1e-12))
check("l1 d5", np.allclose(eng.l1_depth(rng.standard_normal((40, 5))) > -1, True))
check("simplicial enum", (eng.simplicial_counts(P[:30]) == co.simplicial_counts(P[:30])).all())
check("simplicial 3d", (eng.simplicial_counts(rng.standard_normal((12, 3))) >= 0).all())
eng.set_option(E.OPT_SIMPLICIAL_IMPL, E.SIMPLICIAL_COUNT)
check("simplicial count", (eng.simplicial_counts(P, None, 0.0) == co.simplicial_counts(P, None, 0.0)).all())
F = rng.standard_normal((14, 5, 2)).cumsum(1)
check("simplex relax count", (eng.simplex_depth_counts(F, None, True, 0.0) == co.simplex_depth_counts(F, None, True, 0.0)).all())
eng.set_option(E.OPT_SIMPLICIAL_IMPL, E.SIMPLICIAL_AUTO)
check("simplex strict", (eng.simplex_depth_counts(F, None, False) == co.simplex_depth_counts(F, None, False)).all())
from scipy.spatial import ConvexHull
check("oja", np.allclose(eng.oja(P[:25], ConvexHull(P[:25]).volume), co.oja(P[:25], ConvexHull(P[:25]).volume), rtol=1e-12))
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
