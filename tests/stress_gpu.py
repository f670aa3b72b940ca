"""Manual randomized stress run (not collected by pytest): python tests/stress_gpu.py  on a GPU box, from the repo root.
800 band-depth cases (MBD counts + ranks, strict BD through AUTO and the matcher, query subsets) and 120 wide-row
MBD cases against the CPU oracle; prints the number of mismatches (round 1: 0 of 920)."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from statdepth_b200._engine import get_engine, BD_AUTO, BD_MATCH, OPT_BD_IMPL
from oracle import cpu_oracle as oracle
from test_gpu_parity import _random_matrix
oracle.build()
eng = get_engine()
t0 = time.time(); bad = 0; cases = 0
for seed in range(400):
    rng = np.random.default_rng(50_000 + seed)
    n = int(rng.choice([3, 5, 33, 257, 1023, 1024, 1025, 2047, 2049, 5000, 12_000, 30_000]))
    T = int(rng.integers(1, 9))
    X = _random_matrix(rng, T, n)
    if rng.random() < 0.2:
        X[:, int(rng.integers(0, n))] = 1e12 * rng.choice([-1, 1])
    want, wb, wa = oracle.mbd_counts_all(X, j=2, want_ranks=True)
    got = eng.band_depth_counts(X, None, 2, True)
    b, a = eng.band_ranks(X)
    ok = (got == want).all() and (b == wb).all() and (a == wa).all()
    cases += 1
    if not ok:
        bad += 1; print("MBD MISMATCH seed", seed, n, T)
for seed in range(200):
    rng = np.random.default_rng(90_000 + seed)
    n = int(rng.choice([3, 4, 63, 64, 65, 200, 513, 900, 1500]))
    T = int(rng.choice([1, 2, 15, 31, 32, 33, 63, 64, 65, 129, 300]))
    X = _random_matrix(rng, T, n)
    want = oracle.bd_counts(X)
    for impl in (BD_AUTO, BD_MATCH):
        eng.set_option(OPT_BD_IMPL, impl)
        got = eng.band_depth_counts(X, None, 2, False)
        cases += 1
        if not (got == want).all():
            bad += 1; print("BD MISMATCH seed", seed, n, T, impl)
    q = rng.choice(n, size=min(n, 70), replace=False)
    got = eng.band_depth_counts(X, q, 2, False)
    if not (got == want[q]).all():
        bad += 1; print("BD subset MISMATCH seed", seed, n, T)
eng.set_option(OPT_BD_IMPL, BD_AUTO)
for seed in range(60):   # wide rows
    rng = np.random.default_rng(70_000 + seed)
    n = int(rng.choice([60_000, 100_000, 131_072, 131_073, 250_000]))
    T = int(rng.integers(1, 4))
    X = _random_matrix(rng, T, n)
    for j in (2, 3):
        want = oracle.mbd_counts_all(X, j=j)
        got = eng.band_depth_counts(X, None, j, True)
        cases += 1
        if not (got == want).all():
            bad += 1; print("MBD wide MISMATCH seed", seed, n, T, j, eng.timings())
print("cases", cases, "bad", bad, "seconds %.1f" % (time.time() - t0))
