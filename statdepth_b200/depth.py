"""Public API: FunctionalDepth / PointcloudDepth with the reference's signatures and result types.

Mirrors statdepth/depth/depth.py (factories :347-402, result wrappers :14-344) and
statdepth/depth/abstract.py.  The depth values come from the B200 engine; everything below is thin
pandas glue.  plotly is imported lazily inside the plot methods (the reference imports it at
module top, depth.py:4, which makes `import statdepth` fail where plotly is not installed).
"""
from abc import ABC, abstractmethod
from typing import List, Union

import pandas as pd

from ._functional import _functionaldepth, _samplefunctionaldepth
from ._helper import DepthDegeneracy  # noqa: F401  (re-exported like the reference's `from ._helper import *`)
from ._pointcloud import _pointwisedepth, _samplepointwisedepth

__all__ = ['FunctionalDepth', 'PointcloudDepth', 'AbstractDepth']


def _go():
    try:
        import plotly.graph_objects as go
    except ImportError as exc:  # pragma: no cover - plotting is host-only convenience
        raise ImportError('the plot_* helpers need plotly (pip install plotly)') from exc
    return go


class AbstractDepth(ABC):
    """Interface every depth result implements (abstract.py:3-18)."""

    @abstractmethod
    def ordered(self, ascending=False):
        raise NotImplementedError

    @abstractmethod
    def deepest(self, n=1):
        raise NotImplementedError

    @abstractmethod
    def outlying(self, n=1):
        raise NotImplementedError


class _FunctionalDepthSeries(AbstractDepth, pd.Series):
    """pd.Series of depths that remembers the data it was computed from (depth.py:14-65)."""

    def __init__(self, df: pd.DataFrame, depths: pd.Series):
        super().__init__(data=depths)
        self._orig_data = df
        self._depths = depths
        self._ordered_depths = None

    def ordered(self, ascending=False) -> pd.Series:
        """Curves sorted from deepest to most outlying."""
        if self._ordered_depths is None:
            self._ordered_depths = self._depths.sort_values(ascending=ascending)
        return self._ordered_depths

    def _desc(self) -> pd.Series:
        if self._ordered_depths is None:
            self._ordered_depths = self._depths.sort_values(ascending=False)
        return self._ordered_depths

    def deepest(self, n=1) -> pd.Series:
        o = self._desc()
        if n == 1:
            return pd.Series(index=[list(o.index)[0]], data=[o.values[0]])
        return pd.Series(index=o.index[0:n], data=o.values[0:n])

    def outlying(self, n=1) -> pd.Series:
        o = self._desc()
        if n == 1:
            return pd.Series(index=[list(o.index)[-1]], data=[o.values[-1]])
        return pd.Series(index=o.index[-n:], data=o.values[-n:])

    def sorted(self, ascending=False):
        return self.ordered(ascending=ascending)

    def median(self):
        return self.deepest(n=1)

    def quartile(self, ratio=0.5):
        # the reference ignores `ratio` and always takes the lower half (depth.py:55-56); kept
        return self._depths.sort_values().head(int(self._depths.shape[0] * 0.5))

    def get_depths(self):
        return self._depths

    def get_data(self):
        return self._orig_data

    def depths(self):
        return self.get_depths()


class _FunctionalDepthMultivariateDataFrame(AbstractDepth, pd.DataFrame):
    """Placeholder result type of the reference (depth.py:67-83); never produced by the exact path."""

    def __init__(self, depths: pd.DataFrame):
        super().__init__(depths)
        self._depths = depths

    def ordered(self, ascending=False):
        pass

    def deepest(self, n=1):
        pass

    def outlying(self, n=1):
        pass


class _FunctionalDepthUnivariate(_FunctionalDepthSeries):
    """Real-valued curves: samples are COLUMNS of the original frame (depth.py:87-185)."""

    def __init__(self, df: pd.DataFrame, depths: pd.Series):
        super().__init__(df=df, depths=depths)

    def _plot(self, deep_or_outlying: pd.Series, title, xaxis_title, yaxis_title, return_plot, showlegend):
        go = _go()
        cols = self._orig_data.columns
        x = self._orig_data.index
        traces = [go.Scatter(x=x, y=self._orig_data.loc[:, y], mode='lines', name=y,
                             line=dict(color='#6ea8ff', width=.5))
                  for y in cols.difference(deep_or_outlying.index)]
        traces.extend(go.Scatter(x=x, y=self._orig_data.loc[:, y], mode='lines', name=y,
                                 line=dict(color='Red', width=1)) for y in deep_or_outlying.index)
        fig = go.Figure(data=traces, layout=go.Layout(
            title=dict(text=title, y=0.9, x=0.5, xanchor='center', yanchor='top'),
            xaxis=dict(title=xaxis_title), yaxis=dict(title=yaxis_title)))
        fig.update_layout(showlegend=showlegend)
        return fig if return_plot else fig.show()

    def plot_deepest(self, n=1, title=None, xaxis_title=None, yaxis_title=None, return_plot=False,
                     showlegend=False):
        """All curves in blue, the n deepest in red."""
        return self._plot(self.deepest(n=n), title, xaxis_title, yaxis_title, return_plot, showlegend)

    def plot_outlying(self, n=1, title=None, xaxis_title=None, yaxis_title=None, return_plot=False,
                      showlegend=False):
        """All curves in blue, the n most outlying in red."""
        return self._plot(self.outlying(n=n), title, xaxis_title, yaxis_title, return_plot, showlegend)

    def drop_outlying_data(self, n=1) -> pd.DataFrame:
        return self._orig_data.drop(self.outlying(n=n).index, axis=1)

    def get_deepest_data(self, n=1) -> pd.DataFrame:
        return self._orig_data.loc[:, self.deepest(n=n).index]

    def get_outlying_data(self, n=1) -> pd.DataFrame:
        return self._orig_data.loc[:, self.outlying(n=n).index]


class _PointwiseDepth(_FunctionalDepthSeries):
    """Depth of every point of a cloud: samples are ROWS of the original frame (depth.py:187-344)."""

    def __init__(self, df: pd.DataFrame, depths: pd.Series):
        super().__init__(df=df, depths=depths)

    def _scatter(self, go, frame, **marker):
        cols = self._orig_data.columns
        if len(cols) == 3:
            return go.Scatter3d(x=frame[cols[0]], y=frame[cols[1]], z=frame[cols[2]], mode='markers', **marker)
        return go.Scatter(x=frame[cols[0]], y=frame[cols[1]], mode='markers', **marker)

    def plot_depths(self, invert_colors=False, marker=None, return_plot=False, title='', xaxis_title=None,
                    yaxis_title=None):
        go = _go()
        d = 1 - self._depths if invert_colors else self._depths
        ncol = len(self._orig_data.columns)
        if marker is None:
            marker = dict(color=d, colorscale='viridis', size=7)
        if ncol > 3:
            return self._plot_parallel_axis()
        if ncol < 2:
            raise ValueError(f'Error: Dimensionality of data must be >=2. Value found is {ncol}')
        fig = go.Figure(data=[self._scatter(go, self._orig_data, marker=marker)],
                        layout=go.Layout(title=title, xaxis_title=xaxis_title, yaxis_title=yaxis_title))
        fig.update_layout(showlegend=False)
        return fig if return_plot else fig.show()

    def _plot_parallel_axis(self) -> None:
        pass

    def _plot(self, deep_or_outlying: pd.Series, return_plot, title, xaxis_title, yaxis_title):
        go = _go()
        ncol = len(self._orig_data.columns)
        select = self._orig_data.loc[deep_or_outlying.index, :]
        if ncol > 3:
            return self._plot_parallel_axis()
        if ncol < 2:
            raise ValueError(f'Error: Dimensionality of data must be >=2. Value found is {ncol}')
        fig = go.Figure(data=[self._scatter(go, self._orig_data, marker_color='blue', name=''),
                              self._scatter(go, select, marker_color='red', name='')],
                        layout=go.Layout(title=title, xaxis_title=xaxis_title, yaxis_title=yaxis_title))
        fig.update_layout(showlegend=False)
        return fig if return_plot else fig.show()

    def plot_deepest(self, n=1, return_plot=False, title='', xaxis_title=None, yaxis_title=None):
        return self._plot(self.deepest(n=n), return_plot, title, xaxis_title, yaxis_title)

    def plot_outlying(self, n=1, return_plot=False, title='', xaxis_title=None, yaxis_title=None):
        return self._plot(self.outlying(n=n), return_plot, title, xaxis_title, yaxis_title)

    def drop_outlying_data(self, n=1) -> pd.DataFrame:
        return self._orig_data.drop(self.outlying(n=n).index, axis=0)

    def get_deepest_data(self, n=1) -> pd.DataFrame:
        return self._orig_data.loc[self.deepest(n=n).index, :]

    def plot_distribution(self, invert_colors=False, marker=None) -> None:
        self.plot_depths(invert_colors, marker)


def PointcloudDepth(
    data: pd.DataFrame,
    to_compute: pd.Index = None,
    K=None,
    containment='simplex',
    quiet=True,
) -> _PointwiseDepth:
    """Depth of the rows of an n x d DataFrame.  containment in 'simplex' | 'l1' | 'oja' | 'mahalanobis'.
    Signature and result type of statdepth.PointcloudDepth (depth.py:347-359)."""
    if K is not None:
        depth = _samplepointwisedepth(data=data, to_compute=to_compute, K=K, containment=containment)
    else:
        depth = _pointwisedepth(data=data, to_compute=to_compute, containment=containment)
    return _PointwiseDepth(df=data, depths=depth)


def FunctionalDepth(
    data: List[pd.DataFrame],
    to_compute=None,
    K=None,
    J=2,
    containment='r2',
    relax=False,
    deep_check=False,
    quiet=True,
) -> Union[_FunctionalDepthSeries, _FunctionalDepthUnivariate, _FunctionalDepthMultivariateDataFrame]:
    """Band depth of curves.  `[df]` (T rows x n curve columns) is the univariate case; a list of N
    DataFrames (T rows x d channels) the multivariate one.  Signature and result types of
    statdepth.FunctionalDepth (depth.py:362-402)."""
    if K is not None:
        depth = _samplefunctionaldepth(data=data, to_compute=to_compute, K=K, J=J, containment=containment,
                                       relax=relax, deep_check=deep_check, quiet=quiet)
    else:
        depth = _functionaldepth(data=data, to_compute=to_compute, J=J, containment=containment, relax=relax,
                                 deep_check=deep_check, quiet=quiet)
    if isinstance(depth, pd.DataFrame):
        return _FunctionalDepthMultivariateDataFrame(depths=depth)
    elif len(data) == 1:
        return _FunctionalDepthUnivariate(df=data[0], depths=depth)
    else:
        return _FunctionalDepthSeries(df=data[0], depths=depth)
