# Point-cloud timings at BASELINE config 5 (run on the GPU box from the repo root: PYTHONPATH=. python profiles/pointcloud_probe.py)
import time
import numpy as np
from statdepth_b200._engine import get_engine

eng = get_engine()
np.random.seed(4)
P = np.random.normal(size=[50_000, 2])
for name, fn in (("simplicial 2-D, 50k points, all queries", lambda: eng.simplicial_counts(P, None, 0.0)),
                 ("L1 depth, 50k points", lambda: eng.l1_depth(P))):
    fn()
    t = time.perf_counter()
    fn()
    print(f"{name}: {time.perf_counter() - t:.4f} s", eng.timings())
