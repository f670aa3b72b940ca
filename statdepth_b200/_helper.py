"""Argument validation and error types of the depth drivers.

Host-side mirror of statdepth/depth/calculations/_helper.py:13,34-107 (`DepthDegeneracy`,
`_handle_depth_errors`): same checks, in the same order, raising the same exception types with
the same messages, so callers written against the reference behave identically.
"""
from inspect import signature
from typing import Callable, List, Union

import numpy as np
import pandas as pd

__all__ = ["DepthDegeneracy", "_handle_depth_errors", "_check_containment"]


class DepthDegeneracy(Exception):
    """Depth is not well defined for this input (e.g. too few curves to form a simplex)."""


class _DepthDegeneracyTypeError(DepthDegeneracy, TypeError):
    """The reference means to raise DepthDegeneracy at _helper.py:92-93 but the f-string that builds
    its message evaluates ``data[0].shape + 2`` (tuple + int) and a TypeError escapes instead.
    This class is both, so code guarding either exception keeps working."""


FUNCTIONAL_CONTAINMENTS = ("r2", "r2_enum", "simplex")


def _handle_depth_errors(data: List[pd.DataFrame], J: int, containment: Union[Callable, str], relax: bool,
                         deep_check: bool) -> None:
    """Reference: _helper.py:34-107 (checks kept in the reference's order)."""
    if not isinstance(data, list):
        raise ValueError('data must be passed as a list.')
    if not isinstance(J, int):
        raise ValueError('J must be an integer.')
    if not (isinstance(containment, str) or isinstance(containment, Callable)):
        raise ValueError('containment must be of type str or Callable.')
    if not isinstance(deep_check, bool):
        raise ValueError('deep_check must be of type bool.')
    if not isinstance(relax, bool):
        raise ValueError('relax must be of type bool')
    if J < 2:
        raise ValueError('Parameter J must be greater than or equal to 2.')
    if len(data) == 0:
        raise ValueError('No data passed.')
    # NB (reference quirk, _helper.py:83): in the univariate case J is compared with len(data[0]),
    # i.e. the number of ROWS (time points), not the number of curves.
    if len(data) == 1 and J >= len(data[0]) or len(data) > 1 and J >= len(data):
        raise ValueError('Parameter J must be less than the number of observations.')
    if len(data) > 1 and containment == 'r2':
        raise ValueError('containment argument \'r2\' is invalid for multivariate data. Use one of '
                         '[\'r2_enum\', \'simplex \'] or a passed containment method. ')
    if isinstance(data, list) and len(data) < data[0].shape[1] + 2 and containment == 'simplex':
        raise _DepthDegeneracyTypeError(
            'Error: Need at least %d functions to form non-degenerate simplices in %d dimensional space. '
            'Only have %d.' % (data[0].shape[1] + 2, data[0].shape[1], len(data)))
    if deep_check:
        indices = []
        for df in data:
            indices.append(df.index)
            df = df.infer_objects()
            for col in df:
                if not np.issubdtype(df[col].dtype, np.number):
                    raise ValueError('DataFrame must only contain numeric dtypes.')
        if not all([all(indices[0] == i) for i in indices]):
            raise ValueError('DataFrames indices must be the same')


def _check_containment(containment: Union[str, Callable]) -> str:
    """Reference: _select_containment / _is_valid_containment (_containment.py:19-43,178-203).

    Returns the built-in containment name.  A user callable is validated exactly like the reference
    does (3 parameters) and then refused: the reference's plug-in contract is a per-subset Python
    callback, which cannot run inside a CUDA kernel, and this engine has no CPU fallback.
    """
    if isinstance(containment, str):
        if containment in FUNCTIONAL_CONTAINMENTS:
            return containment
        raise ValueError(f'containment argument \'{containment}\' is invalid. Use one of [\'r2\', \'r2_enum\', '
                         f'\'simplex \'] or a pass a custom containment function.')
    params = signature(containment).parameters
    if len(params) != 3:
        raise ValueError('Custom containment method has incorrect number of parameters. Expected 3, recieved {}'
                         .format(len(params)))
    raise NotImplementedError('custom containment callables are evaluated per subset in Python by the reference; '
                              'the B200 engine only runs its built-in containments (\'r2\', \'simplex\') and has '
                              'no CPU fallback.')
