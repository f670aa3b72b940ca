#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # ~3 min, single core

The reference's own tests assert types only (tests/test_statdepth.py:27-73); its only numeric
vectors are two doc examples (docs/index.md:20-42, :98-112).  This script therefore records
(a) those two doc examples and (b) outputs of the reference itself on small seeded inputs that
exercise what its fixtures never do (crossing curves, ties, non-degenerate simplices, to_compute,
J=3, K-sampling, homogeneity).  Inputs are stored next to the outputs so the fixture is
self-contained on the GPU box, where the reference does not exist.
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

sd = ref_shim.load()
from statdepth.homogeneity import FunctionalHomogeneity  # noqa: E402
from statdepth.testing import (generate_noisy_multivariate, generate_noisy_pointcloud,  # noqa: E402
                               generate_noisy_univariate)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
cases = []


def f(a):
    return np.asarray(a, dtype=np.float64).tolist()


def add(kind, name, **kw):
    kw.update(kind=kind, name=name)
    cases.append(kw)
    print("  +", kind, name, flush=True)


def functional(name, X, columns=None, index=None, to_compute=None, **kw):
    df = pd.DataFrame(X, columns=columns, index=index)
    res = sd.FunctionalDepth([df], to_compute=to_compute, **kw)
    add("functional", name, X=f(X), columns=list(df.columns), to_compute=to_compute, kwargs=kw,
        index=list(res.index), depths=f(res.values), ordered_index=list(res.ordered().index))


# ---- doc example, docs/index.md:20-42 ------------------------------------------------------
DOC = np.array([[1, 2, 3, 6.0, 9, 8], [2, 4, 4, 7.0, 9, 8], [3, 5, 4, 6.5, 12, 10],
                [2, 6, 2, 6.0, 11, 10], [1, 2, 1, 7.0, 11, 9]], dtype=np.float64)
cols = ["f_%d" % i for i in range(6)]
idx = ["x_%d" % i for i in range(5)]
for J in (2, 3):
    for relax in (False, True):
        functional("doc_table_J%d_%s" % (J, "relax" if relax else "strict"), DOC, cols, idx, J=J, relax=relax)

# ---- seeded crossing random walks / ties / to_compute -----------------------------------------
rng = np.random.default_rng(11)
W = rng.standard_normal((16, 14)).cumsum(0)
for J in (2, 3):
    for relax in (False, True):
        functional("walk_16x14_J%d_%s" % (J, "relax" if relax else "strict"), W, J=J, relax=relax)
Wt = np.round(rng.standard_normal((12, 13)).cumsum(0))
for relax in (False, True):
    functional("ties_12x13_%s" % ("relax" if relax else "strict"), Wt, J=2, relax=relax)
functional("walk_to_compute", W, to_compute=[3, 0, 9], J=2, relax=True)
functional("walk_labels", W[:, :8], columns=list("abcdefgh"), to_compute=["c", "a"], J=2, relax=False)
G = generate_noisy_univariate(seed=4)
functional("generator_default_seed4", G.values, J=2, relax=False)

# ---- BASELINE config 1 shape (200 curves x 100 points), 2 queries each, ~90 s ---------------
t0 = time.time()
X1 = np.random.default_rng(0).standard_normal((100, 200)).cumsum(0)
df1 = pd.DataFrame(X1)
for relax in (False, True):
    res = sd.FunctionalDepth([df1], to_compute=[0, 1], J=2, relax=relax)
    add("functional_seeded", "cfg1_200x100_%s" % ("relax" if relax else "strict"),
        generator="np.random.default_rng(0).standard_normal((100,200)).cumsum(0)", to_compute=[0, 1],
        kwargs=dict(J=2, relax=relax), depths=f(res.values))
print("cfg1 took %.1f s" % (time.time() - t0))

# ---- K-sampled depth (global np.random state, _functional.py:154-186) --------------------------
np.random.seed(123)
dfk = pd.DataFrame(W)
resk = sd.FunctionalDepth([dfk], K=3, J=2, relax=False)
add("functional_K", "walk_K3_seed123", X=f(W), K=3, np_seed=123, kwargs=dict(J=2, relax=False),
    index=[int(i) for i in resk.index], depths=f(resk.values))

# ---- multivariate simplex ------------------------------------------------------------------------
def multivariate(name, F, to_compute=None, **kw):
    data = [pd.DataFrame(F[i]) for i in range(F.shape[0])]
    res = sd.FunctionalDepth(data, to_compute=to_compute, containment="simplex", **kw)
    add("multivariate", name, F=f(F), to_compute=to_compute, kwargs=kw, index=[int(i) for i in res.index],
        depths=f(res.values))


for seed in (0, 1, 2, 3):
    data = generate_noisy_multivariate(seed=seed)  # reference fixture: 100 % degenerate simplices
    multivariate("generator_deg_seed%d" % seed, np.stack([d.values for d in data]), relax=False)
data = generate_noisy_multivariate(num_curves=7, n=6, d=2, seed=5)
multivariate("generator_deg_d2_relax", np.stack([d.values for d in data]), relax=True)
rng = np.random.default_rng(21)
F2 = rng.standard_normal((8, 6, 2)).cumsum(1)
for relax in (False, True):
    multivariate("walk_8x6x2_%s" % ("relax" if relax else "strict"), F2, relax=relax)
multivariate("walk_8x6x2_to_compute", F2, to_compute=[2, 5], relax=True)
F3 = rng.standard_normal((7, 4, 3)).cumsum(1)
multivariate("walk_7x4x3_relax", F3, relax=True)

# ---- point clouds ------------------------------------------------------------------------------------
from scipy.spatial import ConvexHull  # noqa: E402


def cloud(name, P, containment, to_compute=None, index=None):
    df = pd.DataFrame(P, index=index)
    res = sd.PointcloudDepth(df, to_compute=to_compute, containment=containment)
    extra = {}
    if containment == "oja":
        extra["hull_volume"] = float(ConvexHull(P).volume)
    add("pointcloud", name, P=f(P), containment=containment, to_compute=to_compute,
        index=[int(i) for i in (res.index if res.index is not None else range(len(res)))],
        depths=f(res.values), **extra)


DOC_L1 = np.array([[0.873179, 0.828111], [0.368512, 0.024619], [0.927522, 0.348593],
                   [0.481917, 0.748796], [0.980515, 0.954392]])
cloud("doc_l1", DOC_L1, "l1")  # docs/index.md:98-112 prints .703605 .239076 .458779 .456768 .258959
rng = np.random.default_rng(31)
cloud("l1_15x2", rng.standard_normal((15, 2)), "l1")
cloud("l1_12x3", rng.standard_normal((12, 3)), "l1")
cloud("l1_to_compute", rng.standard_normal((15, 2)), "l1", to_compute=[4, 1])
cloud("simplex_10x2", rng.standard_normal((10, 2)), "simplex")
cloud("simplex_9x3", rng.standard_normal((9, 3)), "simplex")
cloud("simplex_gen_seed2", generate_noisy_pointcloud(n=10, d=2, seed=2).values, "simplex")
cloud("oja_10x2", rng.standard_normal((10, 2)), "oja")
cloud("oja_8x3", rng.standard_normal((8, 3)), "oja")

# ---- functional homogeneity p1..p3 (homogeneity.py:65-153) -----------------------------------------
rng = np.random.default_rng(41)
Fh = rng.standard_normal((10, 7)).cumsum(0)
Gh = rng.standard_normal((10, 6)).cumsum(0) + 0.5
for method in ("p1", "p2", "p3"):
    Fd = pd.DataFrame(Fh, columns=["F%d" % i for i in range(7)])
    Gd = pd.DataFrame(Gh, columns=["G%d" % i for i in range(6)])
    h = FunctionalHomogeneity([Fd], [Gd], method=method, quiet=True).homogeneity()
    val = float(np.asarray(h).ravel()[0])
    add("homogeneity", "functional_%s" % method, F=f(Fh), G=f(Gh), method=method, value=val)

meta = dict(generated_by="tests/golden/make_golden.py", python=sys.version.split()[0],
            numpy=np.__version__, pandas=pd.__version__, scipy=__import__("scipy").__version__,
            note="outputs of the unmodified reference (oracle/ref_shim.py shims only)")
with open(OUT, "w") as fh:
    json.dump(dict(meta=meta, cases=cases), fh, indent=0)
print("wrote", OUT, len(cases), "cases")
