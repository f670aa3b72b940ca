#!/usr/bin/env python
"""Generate tests/golden/reference_vectors_large.json: outputs of the UNMODIFIED reference on samples that are
ABOVE the engine's enumerate -> count switch (64 curves / 64 points), where the counting kernels must reproduce
the LP's tolerance band (scipy.optimize.linprog, _containment.py:164-176) instead of exact sign tests.

Run in the build container only (needs /root/reference); every case is one worker process:

    python tests/golden/make_golden_large.py [-j 6]        # ~25 min on 6 cores (4 ms per LP)

Cases
  mv_deg66      generate_noisy_multivariate(num_curves=66, n=2, d=2, seed=0): every simplex is collinear
                (_generating.py:94-96), relaxed, one query                      -> 2 * C(65,3) LPs
  mv_walk66     66 random-walk curves x 2 points, d = 2 (general position), relaxed, one query
  pc_collinear  point cloud of 70 points: 30 on one line through the query, 20 on a second line, 19 random;
                query 0 and a random one                                           -> 2 * C(69,3) LPs
  pc_lattice    70 points of an integer lattice (many exactly collinear triples and on-edge queries), 2 queries
  pc_big        130 points, 60 of them on two lines, one query                   -> C(129,3) = 357 k LPs
Inputs are stored next to the outputs so that the fixture is self-contained on the GPU box.
"""
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors_large.json")


def f(a):
    return np.asarray(a, dtype=np.float64).tolist()


def _sd():
    from oracle import ref_shim
    return ref_shim.load()


def inputs():
    from oracle import ref_shim
    ref_shim.load()
    from statdepth.testing import generate_noisy_multivariate
    cases = {}
    data = generate_noisy_multivariate(num_curves=66, n=2, d=2, seed=0)
    cases["mv_deg66"] = dict(kind="multivariate", F=np.stack([d.values for d in data]), to_compute=[0],
                             kwargs=dict(relax=True))
    rng = np.random.default_rng(66)
    cases["mv_walk66"] = dict(kind="multivariate", F=rng.standard_normal((66, 2, 2)).cumsum(1), to_compute=[5],
                              kwargs=dict(relax=True))
    # collinear structure through the query point 0
    rng = np.random.default_rng(70)
    p0 = np.array([0.25, -0.5])
    d1 = np.array([1.0, 2.0]) / np.sqrt(5.0)
    d2 = np.array([3.0, -1.0]) / np.sqrt(10.0)
    P = np.vstack([p0[None, :], p0 + rng.uniform(-2, 2, 30)[:, None] * d1, p0 + rng.uniform(-2, 2, 20)[:, None] * d2,
                   rng.standard_normal((19, 2))])
    cases["pc_collinear"] = dict(kind="pointcloud", P=P, to_compute=[0, 57])
    g = np.array([[i, j] for i in range(10) for j in range(7)], dtype=np.float64)  # 70 lattice points
    cases["pc_lattice"] = dict(kind="pointcloud", P=g, to_compute=[24, 0])
    rng = np.random.default_rng(130)
    Pb = np.vstack([p0[None, :], p0 + rng.uniform(-3, 3, 35)[:, None] * d1, p0 + rng.uniform(-3, 3, 25)[:, None] * d2,
                    rng.standard_normal((69, 2))])
    cases["pc_big"] = dict(kind="pointcloud", P=Pb, to_compute=[0])
    return cases


def run(item):
    name, c = item
    sd = _sd()
    t0 = time.time()
    if c["kind"] == "multivariate":
        F = c["F"]
        data = [pd.DataFrame(F[i]) for i in range(F.shape[0])]
        res = sd.FunctionalDepth(data, to_compute=c["to_compute"], containment="simplex", **c["kwargs"])
        out = dict(kind="multivariate", name=name, F=f(F), to_compute=c["to_compute"], kwargs=c["kwargs"],
                   index=[int(i) for i in res.index], depths=f(res.values))
    else:
        P = c["P"]
        res = sd.PointcloudDepth(pd.DataFrame(P), to_compute=c["to_compute"], containment="simplex")
        out = dict(kind="pointcloud", name=name, P=f(P), containment="simplex", to_compute=c["to_compute"],
                   index=[int(i) for i in res.index], depths=f(res.values))
    out["seconds"] = round(time.time() - t0, 1)
    print("  +", name, out["seconds"], "s", flush=True)
    return out


if __name__ == "__main__":
    jobs = 6
    if "-j" in sys.argv:
        jobs = int(sys.argv[sys.argv.index("-j") + 1])
    cs = inputs()
    # one query per worker so that the long cases spread over the cores
    items = []
    for name, c in cs.items():
        for q in c["to_compute"]:
            cc = dict(c)
            cc["to_compute"] = [q]
            items.append(("%s_q%d" % (name, q), cc))
    items.sort(key=lambda it: -(it[1].get("P", np.zeros((0, 2))).shape[0]))
    with ProcessPoolExecutor(max_workers=jobs) as ex:
        results = list(ex.map(run, items))
    import scipy
    meta = dict(generated_by="tests/golden/make_golden_large.py", python=sys.version.split()[0], numpy=np.__version__,
                pandas=pd.__version__, scipy=scipy.__version__,
                note="outputs of the unmodified reference (oracle/ref_shim.py shims only), samples above the "
                     "engine's enumerate->count switch")
    with open(OUT, "w") as fh:
        json.dump(dict(meta=meta, cases=results), fh, indent=0)
    print("wrote", OUT, len(results), "cases")
