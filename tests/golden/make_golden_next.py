#!/usr/bin/env python
"""Generate tests/golden/reference_vectors_next.json by running the UNMODIFIED reference on the rows SURVEY 8(f)
calls "next": K-sampled point-cloud depth (global numpy RNG, _pointcloud.py:97-123), point-cloud homogeneity p1..p3
(homogeneity.py:155-201) and Mahalanobis depth (_pointcloud.py:152-174).  Run in the build container only:

    python tests/golden/make_golden_next.py        # ~2 min
"""
import json
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

sd = ref_shim.load()
from statdepth.homogeneity import PointcloudHomogeneity  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors_next.json")
cases = []


def f(a):
    return np.asarray(a, dtype=np.float64).tolist()


rng = np.random.default_rng(51)
P = rng.standard_normal((12, 2))
for containment, K, seed in (("simplex", 2, 7), ("l1", 3, 8), ("simplex", 3, 9), ("oja", 2, 3)):  # sampled Oja is 0: its subsets
    # are drawn from to_compute = [the point] only (_pointcloud.py:182-193)
    np.random.seed(seed)
    res = sd.PointcloudDepth(pd.DataFrame(P), K=K, containment=containment)
    cases.append(dict(kind="pointcloud_K", name="cloud12_%s_K%d_seed%d" % (containment, K, seed), P=f(P), K=K,
                      np_seed=seed, containment=containment, index=[int(i) for i in res.index], depths=f(res.values)))
    print("  + pointcloud_K", containment, K, flush=True)
np.random.seed(10)
res = sd.PointcloudDepth(pd.DataFrame(P), K=2, containment="l1", to_compute=[3, 7])
cases.append(dict(kind="pointcloud_K", name="cloud12_l1_K2_to_compute", P=f(P), K=2, np_seed=10, containment="l1",
                  to_compute=[3, 7], index=[int(i) for i in res.index], depths=f(res.values)))

Fp = rng.standard_normal((8, 2))  # the reference requires len(F) == len(G) (homogeneity.py:204)
Gp = rng.standard_normal((8, 2)) + 0.4
for containment in ("l1", "simplex"):
    for method in ("p1", "p2", "p3"):
        Fd = pd.DataFrame(Fp, index=["F%d" % i for i in range(8)])
        Gd = pd.DataFrame(Gp, index=["G%d" % i for i in range(8)])
        h = PointcloudHomogeneity(Fd, Gd, method=method, containment=containment)
        cases.append(dict(kind="pointcloud_homogeneity", name="cloud_%s_%s" % (containment, method), F=f(Fp), G=f(Gp),
                          method=method, containment=containment, value=float(h.homogeneity()),
                          F_depths=f(h.F_depths().values), G_depths=f(h.G_depths().values)))
        print("  + pointcloud_homogeneity", containment, method, flush=True)

# multivariate functional homogeneity, p1 (homogeneity.py:136-146; its p2 branch passes its arguments positionally into the
# wrong slots and always raises, its p3 branch is `pass`)
from statdepth.homogeneity import FunctionalHomogeneity  # noqa: E402
rmv = np.random.default_rng(3)
Fm = rmv.standard_normal((7, 3, 2))
Gm = rmv.standard_normal((7, 3, 2)) * 0.3  # G's first curve lies deep inside F
for relax in (True, False):
    h = FunctionalHomogeneity([pd.DataFrame(Fm[i]) for i in range(7)], [pd.DataFrame(Gm[i]) for i in range(7)], method="p1",
                              containment="simplex", relax=relax, quiet=True).homogeneity()
    cases.append(dict(kind="functional_homogeneity_mv", name="mv_p1_%s" % ("relax" if relax else "strict"), F=f(Fm), G=f(Gm),
                      method="p1", relax=relax, value=float(h)))
    print("  + functional_homogeneity_mv p1", relax, float(h), flush=True)

M = np.random.default_rng(52).standard_normal((5, 5))  # n == p: the covariance is singular, the inverse what LAPACK makes of it
res = sd.PointcloudDepth(pd.DataFrame(M), containment="mahalanobis")
cases.append(dict(kind="mahalanobis", name="mahalanobis_5x5", P=f(M), depths=f(res.values)))
res = sd.PointcloudDepth(pd.DataFrame(M), containment="mahalanobis", to_compute=[4, 1])
cases.append(dict(kind="mahalanobis", name="mahalanobis_5x5_to_compute", P=f(M), to_compute=[4, 1], depths=f(res.values)))

meta = dict(generated_by="tests/golden/make_golden_next.py", numpy=np.__version__, pandas=pd.__version__,
            scipy=__import__("scipy").__version__, note="outputs of the unmodified reference (oracle/ref_shim.py shims only)")
with open(OUT, "w") as fh:
    json.dump(dict(meta=meta, cases=cases), fh, indent=0)
print("wrote", OUT, len(cases), "cases")
