"""Shared drivers: run a golden case through a FunctionalDepth / PointcloudDepth implementation."""
import numpy as np
import pandas as pd


def run_functional(case, FunctionalDepth):
    X = np.array(case["X"], dtype=np.float64)
    df = pd.DataFrame(X, columns=case["columns"])
    return FunctionalDepth([df], to_compute=case["to_compute"], **case["kwargs"])


def run_multivariate(case, FunctionalDepth):
    F = np.array(case["F"], dtype=np.float64)
    data = [pd.DataFrame(F[i]) for i in range(F.shape[0])]
    return FunctionalDepth(data, to_compute=case["to_compute"], containment="simplex", **case["kwargs"])


def run_pointcloud(case, PointcloudDepth):
    P = np.array(case["P"], dtype=np.float64)
    return PointcloudDepth(pd.DataFrame(P), to_compute=case["to_compute"], containment=case["containment"])


def check_case(case, FunctionalDepth, PointcloudDepth, rtol=1e-12):
    kind = case["kind"]
    if kind == "functional":
        res = run_functional(case, FunctionalDepth)
        assert list(res.index) == case["index"]
        relax = case["kwargs"].get("relax", False)
        if relax:
            np.testing.assert_allclose(res.values, case["depths"], rtol=rtol, atol=1e-15)
        else:  # strict: integer count / binom -> bit-exact, and so is the ordering
            assert res.values.tolist() == case["depths"]
            assert list(res.ordered().index) == case["ordered_index"]
    elif kind == "multivariate":
        res = run_multivariate(case, FunctionalDepth)
        assert [int(i) for i in res.index] == case["index"]
        np.testing.assert_allclose(res.values, case["depths"], rtol=rtol, atol=1e-15)
    elif kind == "pointcloud":
        res = run_pointcloud(case, PointcloudDepth)
        np.testing.assert_allclose(res.values, case["depths"], rtol=rtol, atol=1e-15)
        if case["containment"] != "oja" or case["to_compute"] is not None:
            assert [int(i) for i in res.index] == case["index"]
    else:
        raise AssertionError(kind)
