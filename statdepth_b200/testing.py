"""Synthetic data generators with the reference's signatures (statdepth/testing/_generating.py).

Kept so that scripts written against `statdepth.testing` run unchanged.  The random streams are
the reference's (np.random.seed + the same draws in the same order).  Note what these fixtures are:
every univariate / multivariate "noisy" curve is a scalar multiple of one base curve
(_generating.py:42-44, 94-96), so curves never cross and every simplex is degenerate.
"""
from typing import List, Union

import numpy as np
import pandas as pd

__all__ = ['generate_noisy_univariate', 'generate_noisy_multivariate', 'generate_noisy_pointcloud']


def generate_noisy_univariate(data: Union[list, np.ndarray] = None, n: int = 20, columns=None, index=None,
                              seed=None) -> pd.DataFrame:
    """n curves (columns) = data * r_k with r_k ~ U(0,1); _generating.py:5-51."""
    np.random.seed(seed)
    if data is None:
        data = np.random.rand(n)
    cols = {}
    for k in range(n):
        cols[k] = np.multiply(data, np.random.rand())
    df = pd.DataFrame(cols)
    if index is not None:
        df.index = index
    if columns is not None:
        df.columns = columns
    return df


def generate_noisy_multivariate(data: pd.DataFrame = None, num_curves: int = 5, n: int = 10, d: int = 3,
                                columns=None, index=None, seed=None) -> List[pd.DataFrame]:
    """num_curves functions (n rows x d channels) = base * r; _generating.py:53-106."""
    np.random.seed(seed)
    if data is None:
        data = np.random.rand(n, d)
    fs = []
    for _ in range(num_curves):
        fs.append(pd.DataFrame(data) * np.random.rand())
    for df in fs:
        if index is not None:
            df.index = index
        if columns is not None:
            df.columns = columns
    return fs


def generate_noisy_pointcloud(n: int = 50, d: int = 2, columns=None, index=None, seed=None) -> pd.DataFrame:
    """n standard-normal points in R^d; _generating.py:108-140."""
    np.random.seed(seed)
    df = pd.DataFrame(np.random.normal(size=[n, d]))
    if columns is not None:
        df.columns = columns
    if index is not None:
        df.index = index
    return df
