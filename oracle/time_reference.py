#!/usr/bin/env python
"""Time the UNMODIFIED reference (statdepth.FunctionalDepth, pure Python) on this machine's host cores.

TEST / BENCH INFRASTRUCTURE (bench.py's `cpu_baseline.reference_python`); never imported by the product.

    python oracle/time_reference.py [--curves-per-worker 1] [--workers P] [--relax 0|1]

Workload: BASELINE config 1 -- univariate band depth (containment='r2', J=2) of 200 random-walk curves x 100 time
points, `np.random.default_rng(0).standard_normal((100, 200)).cumsum(0)` (the shape the reference can actually run:
~20 s per depth evaluation per core).  The reference is single-threaded and embarrassingly parallel over query
curves, so P worker processes each evaluate `--curves-per-worker` curves through the reference's own
public API, `FunctionalDepth([df], to_compute=[...])`; the result is seconds per depth-eval per core and the
extrapolated time of the full configuration.  Prints one JSON object.  The reference package comes from
/root/reference when present, else from the staged copy oracle/_ref/ (oracle/make_ref.sh).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _work(args):
    cols, relax = args
    import numpy as np
    import pandas as pd
    from oracle import ref_shim
    sd = ref_shim.load()
    X = np.random.default_rng(0).standard_normal((100, 200)).cumsum(0)
    df = pd.DataFrame(X)
    t0 = time.perf_counter()
    res = sd.FunctionalDepth([df], to_compute=list(cols), J=2, relax=bool(relax))
    return time.perf_counter() - t0, [float(v) for v in res.values]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=min(os.cpu_count() or 1, 32))
    ap.add_argument("--curves-per-worker", type=int, default=1)
    ap.add_argument("--relax", type=int, default=0)
    a = ap.parse_args()
    from oracle import ref_shim
    if not ref_shim.available():
        print(json.dumps({"unavailable": "reference package not staged (oracle/make_ref.sh) and /root/reference absent"}))
        return
    import multiprocessing as mp
    jobs = [(list(range(w * a.curves_per_worker, (w + 1) * a.curves_per_worker)), a.relax) for w in range(a.workers)]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(a.workers) as pool:
        out = pool.map(_work, jobs)
    wall = time.perf_counter() - t0
    evals = a.workers * a.curves_per_worker
    busy = sum(t for t, _ in out)
    print(json.dumps({
        "workload": "cfg1: FunctionalDepth([df], containment='r2', J=2, relax=%s), 200 curves x 100 points" % bool(a.relax),
        "api": "statdepth.FunctionalDepth (unmodified reference, %s)" % ref_shim.REFERENCE_ROOT,
        "workers": a.workers, "depth_evals": evals, "wall_s": wall,
        "s_per_depth_eval_per_core": busy / evals,
        "depth_evals_per_s_all_cores": evals / wall,
        "extrapolated_full_cfg1_s_all_cores": 200.0 / (evals / wall),
        "depths_first_worker": out[0][1],
    }))


if __name__ == "__main__":
    main()
