"""Homogeneity coefficients between two samples, driven by the B200 depth engine.

Mirrors statdepth/homogeneity/homogeneity.py: FunctionalHomogeneity / PointcloudHomogeneity
(:9-63), the p1..p3 coefficient definitions (:65-201) and P1_homogeneity / P2_homogeneity
(:214-307).  Documented differences:
  * the caller's F is never mutated (the reference writes a 'g_deepest' column into it, :101,:249);
  * the reference's pointcloud 'p4' recursion is broken (it multiplies tuples, :195-196) and its
    functional 'p4' raises NotImplementedError (:134-135): both raise NotImplementedError here.
`permutation_test` is new (the reference has coefficients but no resampling): it evaluates a
coefficient on B label permutations of the pooled sample.
"""
from typing import List

import numpy as np
import pandas as pd

from .depth import FunctionalDepth, PointcloudDepth

__all__ = ['FunctionalHomogeneity', 'PointcloudHomogeneity', 'P1_homogeneity', 'P2_homogeneity',
           'permutation_test']

_BAD_METHOD = '{} is not a valid depth method for the given data. Use one of [\'p1\', \'p2\', \'p3\', \'p4\']'


def _handle_errors(F, G, method='p1'):
    if len(F) != len(G):
        raise ValueError('F and G must have data of the same length')
    if len(F) == 1 and F[0].shape[0] != G[0].shape[0]:
        raise ValueError('Univariate data must have same number of time indices to check containment.')


def _functionalhomogeneity(F: List[pd.DataFrame], G: List[pd.DataFrame], K=None, J=2, containment='r2',
                           method='p1', relax=False, deep_check=False, quiet=False):
    _handle_errors(F, G, method)
    kw = dict(K=K, J=J, containment=containment, relax=relax, deep_check=deep_check, quiet=quiet)
    G_depths = FunctionalDepth(data=G, **kw)
    if len(F) == 1:
        Fd, Gd = F[0], G[0]
        G_deepest = G_depths.get_deepest_data(n=1)
        Fg = Fd.drop('g_deepest', axis=1) if 'g_deepest' in Fd.columns else Fd.copy()
        Fg.loc[:, 'g_deepest'] = G_deepest.iloc[:, 0].values
        G_deep_in_F = FunctionalDepth([Fg], to_compute=['g_deepest'], **kw)
        Fd = Fg.drop('g_deepest', axis=1)
        if method == 'p1':
            return G_deep_in_F
        elif method == 'p2':
            F_depths = FunctionalDepth([Fd], **kw)
            return np.abs(G_deep_in_F - F_depths.median().iloc[0])
        elif method == 'p3':
            if K is None and containment == 'r2' and J in (2, 3) and not set(Gd.columns) & set(Fd.columns):
                # the |G| single-query depth runs of the reference (homogeneity.py:121-133), each of one curve of G
                # inside F u {that curve}, as ONE batched engine call on the pooled matrix
                t = _depths_of_each_in(Fd, Gd, J, relax)
            else:
                t = []
                for col in Gd.columns:
                    Fc = Fd.copy()
                    Fc.loc[:, col] = Gd.loc[:, col].values
                    t.append(FunctionalDepth([Fc], to_compute=[col], K=K, J=J, containment=containment, relax=relax,
                                             deep_check=deep_check).loc[col])
            depths_G_in_F = pd.Series(index=list(Gd.columns), data=t).sort_values(ascending=False)
            return depths_G_in_F.iloc[0] / G_depths.median().iloc[0]
        elif method == 'p4':
            raise NotImplementedError()
        raise ValueError(_BAD_METHOD.format(method))
    else:
        G_deepest = G[G_depths.index[0]]
        Fx = list(F) + [G_deepest]
        G_deep_in_F = FunctionalDepth(Fx, to_compute=[len(Fx) - 1], K=K, J=J, containment=containment, relax=relax,
                                      deep_check=deep_check).ordered().iloc[0]
        if method == 'p1':
            return G_deep_in_F / G_depths.median().iloc[0]
        elif method == 'p2':
            F_depths = FunctionalDepth(F, K=K, J=J, containment=containment, relax=relax, deep_check=deep_check)
            return 1 - np.abs(G_deep_in_F - F_depths.median().iloc[0])
        elif method == 'p3':
            return None  # the reference's branch is `pass`
        raise ValueError(_BAD_METHOD.format(method))


def _depths_of_each_in(Fd: pd.DataFrame, Gd: pd.DataFrame, J: int, relax: bool) -> np.ndarray:
    """depth of every column g of Gd inside Fd u {g}: |G| sub-populations of the pooled matrix, one batched call
    per subset size j (sd_band_depth_batched_f64)."""
    from scipy.special import binom

    from . import _functional
    eng = _functional.get_engine()
    X = np.ascontiguousarray(np.concatenate([Fd.to_numpy(dtype=np.float64), Gd.to_numpy(dtype=np.float64)], axis=1))
    T, nF, nG = X.shape[0], Fd.shape[1], Gd.shape[1]
    mem = np.zeros((nG, nF + nG), dtype=np.uint8)
    mem[:, :nF] = 1
    mem[np.arange(nG), nF + np.arange(nG)] = 1
    q = (nF + np.arange(nG, dtype=np.int64))[:, None]
    depth = np.zeros(nG)
    for j in range(2, J + 1):
        s_nj = eng.band_depth_counts_batched(X, mem, q, j, relax)[:, 0].astype(np.float64)
        if relax:
            s_nj = s_nj / float(T)
        depth = depth + s_nj / binom(nF + 1, j)
    return depth


def _pointcloudhomogeneity(F: pd.DataFrame, G: pd.DataFrame, K=None, containment='simplex', method='p1'):
    _handle_errors(F, G, method)
    G_depths = PointcloudDepth(data=G, K=K, containment=containment)
    F_depths = PointcloudDepth(data=F, K=K, containment=containment)
    G_deepest = G_depths.get_deepest_data(n=1).copy()
    G_deepest.index = ['g_deepest']
    Fg = pd.concat([F, G_deepest])
    G_deep_in_F = PointcloudDepth(Fg, to_compute=['g_deepest'], K=K, containment=containment) \
        .ordered().loc['g_deepest']
    if method == 'p1':
        hom = G_deep_in_F / F_depths.median().iloc[0]
    elif method == 'p2':
        hom = 1 - np.abs(G_deep_in_F - F_depths.median().iloc[0])
    elif method == 'p3':
        t = []
        for point in G.index:
            Fp = F.copy()
            Fp.loc[point, :] = G.loc[point, :]
            t.append(PointcloudDepth(Fp, to_compute=[point], K=K, containment=containment).loc[point])
        depths_G_in_F = pd.Series(index=list(G.index), data=t).sort_values(ascending=False)
        hom = depths_G_in_F.iloc[0] / G_depths.median().iloc[0]
    elif method == 'p4':
        raise NotImplementedError()
    else:
        raise ValueError(_BAD_METHOD.format(method))
    return F_depths, G_depths, hom


class FunctionalHomogeneity:
    def __init__(self, F, G, method='p1', K=None, J=2, containment='r2', relax=False, deep_check=False,
                 quiet=False):
        self._orig_F = F
        self._orig_G = G
        self._hom = _functionalhomogeneity(F=F, G=G, K=K, J=J, containment=containment, method=method,
                                           relax=relax, deep_check=deep_check, quiet=quiet)

    def homogeneity(self):
        return self._hom

    def __str__(self):
        return str(self.homogeneity())

    def __repr__(self):
        return str(self.homogeneity())


class PointcloudHomogeneity:
    def __init__(self, F, G, method='p1', K=None, J=None, containment='simplex', relax=False, deep_check=False):
        self._orig_F = F
        self._orig_G = G
        self._F_depths, self._G_depths, self._hom = _pointcloudhomogeneity(F=F, G=G, K=K,
                                                                           containment=containment, method=method)

    def F_depths(self):
        return self._F_depths

    def G_depths(self):
        return self._G_depths

    def homogeneity(self):
        return self._hom

    def __str__(self):
        return str(self.homogeneity())

    def __repr__(self):
        return str(self.homogeneity())


def P1_homogeneity(F: pd.DataFrame, G: pd.DataFrame, K=None, J=2, containment='r2', relax=False,
                   quiet=False) -> float:
    """Depth, inside F, of the deepest curve of G (closer to the top depth of F = more homogeneous)."""
    G_depth = FunctionalDepth(data=[G], K=K, J=J, containment=containment, relax=relax, quiet=quiet)
    G_deepest = G_depth.get_deepest_data()
    Fg = F.copy()
    Fg.loc[:, 'G_deepest'] = G_deepest.iloc[:, 0].values
    G_deep_in_F = FunctionalDepth([Fg], to_compute=['G_deepest'], K=K, J=J, containment=containment, relax=relax,
                                  quiet=quiet)
    return G_deep_in_F.iloc[0]


def P2_homogeneity(F: pd.DataFrame, G: pd.DataFrame, K=None, J=2, containment='r2', relax=False,
                   quiet=False) -> float:
    """|P1(F, G) - depth of F's own deepest curve| (closer to 0 = more homogeneous).

    The reference computes the second term AFTER P1_homogeneity added 'G_deepest' to the caller's F
    (homogeneity.py:249, 298), i.e. on F u {g}; that is reproduced here on a copy."""
    P1_F_G = P1_homogeneity(F=F, G=G, K=K, J=J, containment=containment, relax=relax, quiet=quiet)
    G_depth = FunctionalDepth(data=[G], K=K, J=J, containment=containment, relax=relax, quiet=quiet)
    Fg = F.copy()
    Fg.loc[:, 'G_deepest'] = G_depth.get_deepest_data().iloc[:, 0].values
    P1_F_F = FunctionalDepth(data=[Fg], K=K, J=J, containment=containment, relax=relax, quiet=quiet) \
        .deepest().iloc[0]
    return np.abs(P1_F_G - P1_F_F)


def _perm_stats_batched(pooled: pd.DataFrame, nF: int, perms: np.ndarray, method: str, relax: bool) -> np.ndarray:
    """p1 / p2 / p3 statistics of all permutations with two or three BATCHED engine calls
    (sd_band_depth_batched_f64: one launch sequence for all permutations) instead of 2-3 calls per
    permutation.  Same arithmetic and the same tie-breaking (pandas sort) as FunctionalHomogeneity."""
    from scipy.special import binom

    from ._engine import get_engine
    eng = get_engine()
    X = np.ascontiguousarray(pooled.to_numpy(dtype=np.float64))
    T, n = X.shape
    B = perms.shape[0]
    nG = n - nF

    def depths(membership, queries, sizes):
        cnt = eng.band_depth_counts_batched(X, membership, queries, 2, relax).astype(np.float64)
        if relax:
            cnt = cnt / float(T)
        return cnt / binom(sizes, 2)[:, None]

    rows = np.arange(B)[:, None]
    # (1) depth of every curve of G_b inside G_b  -> deepest curve of G_b
    memG = np.zeros((B, n), dtype=np.uint8)
    memG[rows, perms[:, nF:]] = 1
    qG = np.ascontiguousarray(perms[:, nF:])
    dG = depths(memG, qG, np.full(B, nG))
    # deepest = first label of sort_values(ascending=False); without a tie at the maximum that is the argmax,
    # with one the pandas call itself decides (its sort is not stable)
    deepest = qG[np.arange(B), dG.argmax(axis=1)]
    tied = np.flatnonzero((dG == dG.max(axis=1, keepdims=True)).sum(axis=1) > 1)
    for b in tied:
        deepest[b] = pd.Series(index=qG[b], data=dG[b]).sort_values(ascending=False).index[0]
    # (2) depth of that curve inside F_b u {g}
    memF = np.zeros((B, n), dtype=np.uint8)
    memF[rows, perms[:, :nF]] = 1
    memF[np.arange(B), deepest] = 1
    g_in_F = depths(memF, deepest[:, None], np.full(B, nF + 1))[:, 0]
    if method == 'p1':
        return g_in_F
    if method == 'p3':
        # max over g in G_b of depth(g in F_b u {g}) / top depth of G_b in G_b: B * nG sub-populations, one call
        memE = np.zeros((B * nG, n), dtype=np.uint8)
        rowsE = np.arange(B * nG)
        memE[rowsE[:, None], np.repeat(perms[:, :nF], nG, axis=0)] = 1  # row b * nG + k: F_b ...
        gE = qG.reshape(-1)                                              # ... plus the k-th curve of G_b
        memE[rowsE, gE] = 1
        dE = depths(memE, gE[:, None], np.full(B * nG, nF + 1))[:, 0].reshape(B, nG)
        return dE.max(axis=1) / dG.max(axis=1)  # `.median()` of the reference's result types is the DEEPEST curve
    # (3) p2: | depth(g in F u {g}) - max depth of F_b in F_b |
    memF0 = np.zeros((B, n), dtype=np.uint8)
    memF0[rows, perms[:, :nF]] = 1
    dF = depths(memF0, np.ascontiguousarray(perms[:, :nF]), np.full(B, nF))
    return np.abs(g_in_F - dF.max(axis=1))


def permutation_test(F: pd.DataFrame, G: pd.DataFrame, method='p1', B=200, seed=None, J=2, containment='r2',
                     relax=True, batched=True) -> dict:
    """Permutation null of a functional homogeneity coefficient (NEW: no reference counterpart).

    The observed statistic is FunctionalHomogeneity([F], [G], method); the null re-labels the pooled
    curves B times with np.random.default_rng(seed).permutation and re-evaluates it.  Permutations are
    independent: with torch.distributed initialised they are split across ranks and all-gathered; within a
    rank p1 / p2 are evaluated for all permutations by batched engine calls (`batched=False` loops instead).
    Returns dict(observed, null (B floats), p_value = P(null more extreme than observed)).
    """
    from . import _dist

    def stat(Fd, Gd):
        h = FunctionalHomogeneity([Fd], [Gd], method=method, J=J, containment=containment, relax=relax,
                                  quiet=True).homogeneity()
        return float(np.asarray(h).ravel()[0])

    pooled = pd.concat([F, G], axis=1)
    pooled.columns = range(pooled.shape[1])
    nF = F.shape[1]
    rng = np.random.default_rng(seed)
    perms = np.stack([rng.permutation(pooled.shape[1]) for _ in range(B)]) if B > 0 else np.zeros((0, 0), int)
    observed = stat(pooled.iloc[:, :nF], pooled.iloc[:, nF:])

    def run(block):
        block = list(block)
        if batched and method in ('p1', 'p2', 'p3') and J == 2 and containment == 'r2' and block:
            return _perm_stats_batched(pooled, nF, perms[block], method, relax)
        return np.array([stat(pooled.iloc[:, perms[b][:nF]], pooled.iloc[:, perms[b][nF:]]) for b in block])

    # the inner depth calls must not shard again while permutations are sharded across ranks
    rank, size = _dist.world()
    if size > 1:
        lo, hi = _dist.block(B, rank, size)
        with _dist.local_only():
            local = run(range(lo, hi))
        null = _dist.allgather_blocks(local.astype(np.float64), B)
    else:
        null = run(range(B))
    if method == 'p2':   # p2: small = homogeneous
        p = float((np.sum(null >= observed) + 1) / (B + 1))
    else:                # p1 / p3: large = homogeneous
        p = float((np.sum(null <= observed) + 1) / (B + 1))
    return dict(observed=observed, null=null, p_value=p)
