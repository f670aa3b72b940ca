# Homogeneity permutation test of BASELINE config 5 (run on the GPU box from the repo root: PYTHONPATH=. python profiles/perm_probe.py)
import time
import numpy as np, pandas as pd
from statdepth_b200.homogeneity import permutation_test
rng = np.random.default_rng(5)
F = pd.DataFrame(rng.standard_normal((256, 256)).cumsum(0))
G = pd.DataFrame(rng.standard_normal((256, 256)).cumsum(0) + 0.5)
for relax in (True, False):
    for method in ("p1", "p2"):
        permutation_test(F, G, method=method, B=20, seed=5, relax=relax)
        t = time.perf_counter()
        out = permutation_test(F, G, method=method, B=1000, seed=5, relax=relax)
        print(f"permutation test {method}, 2 x 256 curves x 256 points, 1000 permutations, relax={relax}: "
              f"{time.perf_counter() - t:.3f} s  p={out['p_value']:.4f}")

# the batched engine call alone: 1000 sub-populations of 256 of the 512 pooled curves, every member a query
from statdepth_b200._engine import get_engine
eng = get_engine()
X = np.ascontiguousarray(pd.concat([F, G], axis=1).to_numpy())
perms = np.stack([rng.permutation(512) for _ in range(1000)])
mem = np.zeros((1000, 512), dtype=np.uint8)
for b in range(1000):
    mem[b, perms[b, :256]] = 1
qs = np.ascontiguousarray(perms[:, :256])
for relax in (True, False):
    eng.band_depth_counts_batched(X, mem, qs, 2, relax)
    t = time.perf_counter()
    eng.band_depth_counts_batched(X, mem, qs, 2, relax)
    print(f"sd_band_depth_batched_f64, 1000 x (256 curves x 256 points), 256 queries each, relax={relax}: "
          f"{time.perf_counter() - t:.4f} s", eng.timings())
