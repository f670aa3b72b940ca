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

# BASELINE config 4 (relaxed, 2 channels): 5 000 curves x 256 time points, all queries
F = np.random.default_rng(3).standard_normal((5000, 256, 2)).cumsum(1)
eng.set_option(3, 1)  # SD_OPT_PROFILE
for _ in range(2):
    t = time.perf_counter()
    c = eng.simplex_depth_counts(F, None, True, 0.0)
    dt = time.perf_counter() - t
print(f"relaxed simplex depth, 5000 curves x 256 points x 2 channels: {dt:.4f} s", eng.timings(),
      {k: round(v / 1e6, 2) for k, v in eng.phase_ns().items()})
