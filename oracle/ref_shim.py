"""Import shim for the UNMODIFIED reference package (test infrastructure only).

The reference (``/root/reference``, pure Python) cannot be imported as-is in this
image for two reasons that have nothing to do with its arithmetic:

* ``statdepth/depth/depth.py:4`` imports ``plotly.graph_objects`` at module top and
  plotly is not installed -> empty stub modules are pre-seeded in ``sys.modules``;
* pandas here is 3.x, ``DataFrame.append`` is gone, and the reference calls it in
  ``_pointcloud.py:118,198`` and ``homogeneity.py:173`` -> a concat-based shim.

Nothing under ``/root/reference`` is modified or copied.  This module is used ONLY by
``tests/golden/make_golden.py`` (fixture generation, run in the build container) and
by optional live-reference tests that skip when ``/root/reference`` is absent (it does
not exist on the GPU box).  Product code never imports it.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")  # oracle/make_ref.sh (travels to the GPU box)
REFERENCE_ROOT = os.environ.get("STATDEPTH_REFERENCE",
                                "/root/reference" if os.path.isdir("/root/reference/statdepth") else _STAGED)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "statdepth"))


def load():
    """Return the reference ``statdepth`` package (raises ImportError if absent)."""
    if not available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    import pandas as pd

    if "plotly" not in sys.modules:
        plotly = types.ModuleType("plotly")
        go = types.ModuleType("plotly.graph_objects")
        plotly.graph_objects = go
        sys.modules["plotly"] = plotly
        sys.modules["plotly.graph_objects"] = go
    if not hasattr(pd.DataFrame, "append"):
        def _append(self, other, ignore_index=False):
            if isinstance(other, pd.Series):
                other = other.to_frame().T
            return pd.concat([self, other], ignore_index=ignore_index)
        pd.DataFrame.append = _append
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import statdepth  # noqa: E402

    return statdepth
