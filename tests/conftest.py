import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_vectors.json")) as fh:
        doc = json.load(fh)
    return {c["name"]: c for c in doc["cases"]}


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle


@pytest.fixture(scope="session")
def engine():
    """The real CUDA engine.  GPU tests fail loudly (not skip) if the extension cannot start."""
    from statdepth_b200 import get_engine
    return get_engine()


class OracleBackedEngine:
    """TEST DOUBLE with the Engine interface, answering from the CPU oracle.

    Lets the CPU test-suite exercise the Python host layer (label handling, K-sampling replay,
    float assembly, result types, distributed sharding) without a GPU.  Lives in tests/ only; the
    product never routes through it (statdepth_b200 raises EngineUnavailable without a GPU)."""

    def __init__(self):
        from oracle import cpu_oracle
        self.o = cpu_oracle

    def band_depth_counts(self, X, queries=None, j=2, relax=False):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if relax:
            allc = self.o.mbd_counts_all(X, j=j)
            return allc if queries is None else allc[np.asarray(queries, dtype=np.int64)]
        return self.o.bd_counts(X, queries, j=j)

    def band_ranks(self, X):
        _, b, a = self.o.mbd_counts_all(np.ascontiguousarray(X, dtype=np.float64), want_ranks=True)
        return b, a

    def band_depth_counts_batched(self, X, membership, queries, j=2, relax=False):
        out = np.zeros(queries.shape, dtype=np.int64)
        for b in range(queries.shape[0]):
            cols = np.flatnonzero(membership[b])
            loc = {int(c): i for i, c in enumerate(cols)}
            ql = [loc[int(g)] for g in queries[b]]
            out[b] = self.band_depth_counts(X[:, cols], ql, j, relax)
        return out

    def simplex_depth_counts(self, F, queries=None, relax=False, tol=1e-7):
        return self.o.simplex_depth_counts(F, queries, relax, tol)

    def simplicial_counts(self, P, queries=None, tol=1e-7):
        return self.o.simplicial_counts(P, queries, tol)

    def l1_depth(self, P, queries=None):
        return self.o.l1_depth(P, queries)

    def cloud_blocks(self, P, members, offsets, query_pos, kind, tol=1e-7, hull_volumes=None):
        P = np.ascontiguousarray(P, dtype=np.float64)
        out = np.zeros(len(query_pos))
        for b, qp in enumerate(query_pos):
            sub = np.ascontiguousarray(P[np.asarray(members[offsets[b]:offsets[b + 1]], dtype=np.int64)])
            if kind == "simplex":
                out[b] = self.o.simplicial_counts(sub, [qp], tol)[0]
            elif kind == "l1":
                out[b] = self.o.l1_depth(sub, [qp])[0]
            else:
                out[b] = self.o.oja(sub, hull_volumes[b], [qp])[0]
        return out

    def oja(self, P, hull_volume, queries=None, pool=None):
        return self.o.oja(P, hull_volume, queries, pool)


@pytest.fixture()
def host_on_oracle(monkeypatch):
    """Route statdepth_b200's host layer to the oracle-backed test double (CPU tests only)."""
    fake = OracleBackedEngine()
    import statdepth_b200._functional as f
    import statdepth_b200._pointcloud as p
    monkeypatch.setattr(f, "get_engine", lambda device=None: fake)
    monkeypatch.setattr(p, "get_engine", lambda device=None: fake)
    return fake
