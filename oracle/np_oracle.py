"""Vectorised numpy restatement of the band-depth closed forms.

TEST INFRASTRUCTURE ONLY.  Independent of the C oracle (different code path: numpy sort /
searchsorted / integer matmul), used to cross-check it and as a readable spec:

  others strictly below / above the query at time t:   b_t, a_t
  relaxed numerator (j)  = sum_t [ C(n-1,j) - C(b_t,j) - C(a_t,j) ]       (_containment.py:76,80)
  strict  numerator (2)  = #{pairs (p<r) of others : V[p,r] == 0},  V = Sb Sb^T + Sa Sa^T
  depth_J = sum_{j=2..J} numerator_j / (T if relaxed else 1) / C(n, j)     (_functional.py:229,253)
"""
from math import comb

import numpy as np


def ranks(X):
    """b[t, c], a[t, c]: number of curves strictly below / above curve c at time t.  X is [T, n]."""
    X = np.asarray(X, dtype=np.float64)
    S = np.sort(X, axis=1)
    b = np.empty(X.shape, dtype=np.int64)
    a = np.empty(X.shape, dtype=np.int64)
    n = X.shape[1]
    for t in range(X.shape[0]):
        b[t] = np.searchsorted(S[t], X[t], side="left")
        a[t] = n - np.searchsorted(S[t], X[t], side="right")
    return b, a


def _c(m, j):
    m = np.asarray(m, dtype=object)
    return np.vectorize(lambda v: comb(int(v), j), otypes=[object])(m)


def mbd_counts(X, j=2):
    """Relaxed numerator for every curve, exact Python ints (object array)."""
    b, a = ranks(X)
    n = X.shape[1]
    full = comb(n - 1, j)
    if j == 2:
        term = full - b * (b - 1) // 2 - a * (a - 1) // 2
        return term.sum(axis=0)
    return (full - _c(b, j) - _c(a, j)).sum(axis=0)


def bd_counts(X, queries=None):
    """Strict J=2 numerator via the violation Gram (dense int matmul): small n only."""
    X = np.asarray(X, dtype=np.float64)
    T, n = X.shape
    qs = range(n) if queries is None else queries
    out = []
    for c in qs:
        others = [k for k in range(n) if k != c]
        Sb = (X[:, others] < X[:, [c]]).T.astype(np.int32)
        Sa = (X[:, others] > X[:, [c]]).T.astype(np.int32)
        V = Sb @ Sb.T + Sa @ Sa.T
        iu = np.triu_indices(len(others), k=1)
        out.append(int((V[iu] == 0).sum()))
    return np.array(out, dtype=np.int64)


def depth_from_counts(counts_by_j, n, T, relax):
    """Float64 depth exactly as the host computes it: sum_j (count_j [/T]) / C(n,j)."""
    d = np.zeros(len(next(iter(counts_by_j.values()))), dtype=np.float64)
    for j, cnt in counts_by_j.items():
        num = np.asarray([float(int(v)) for v in cnt], dtype=np.float64)
        if relax:
            num = num / float(T)
        d = d + num / float(comb(n, j))
    return d


def l1_depth(P):
    """Vectorised _L1_depth (_pointcloud.py:125-150); summation order differs from the reference."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    out = np.empty(n)
    for p in range(n):
        diff = np.delete(P, p, axis=0) - P[p]
        nrm = np.linalg.norm(diff, axis=1, keepdims=True)
        out[p] = 1.0 - np.linalg.norm((diff / nrm).sum(axis=0)) / n
    return out
