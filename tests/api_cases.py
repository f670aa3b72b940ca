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


def load_next_cases():
    """tests/golden/reference_vectors_next.json: K-sampled point clouds, point-cloud homogeneity, Mahalanobis."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors_next.json")) as fh:
        return json.load(fh)["cases"]


def check_next_case(case, PointcloudDepth, PointcloudHomogeneity, rtol=1e-12):
    """SURVEY 8(f) rows against the unmodified reference's outputs."""
    kind = case["kind"]
    if kind == "pointcloud_K":
        np.random.seed(case["np_seed"])  # the reference samples blocks from the GLOBAL numpy RNG (_pointcloud.py:114)
        res = PointcloudDepth(pd.DataFrame(np.array(case["P"])), K=case["K"], containment=case["containment"],
                              to_compute=case.get("to_compute"))
        assert [int(i) for i in res.index] == case["index"]
        np.testing.assert_allclose(res.values, case["depths"], rtol=rtol, atol=1e-15)
    elif kind == "pointcloud_homogeneity":
        F = pd.DataFrame(np.array(case["F"]), index=["F%d" % i for i in range(len(case["F"]))])
        G = pd.DataFrame(np.array(case["G"]), index=["G%d" % i for i in range(len(case["G"]))])
        h = PointcloudHomogeneity(F, G, method=case["method"], containment=case["containment"])
        np.testing.assert_allclose(float(h.homogeneity()), case["value"], rtol=rtol)
        np.testing.assert_allclose(h.F_depths().values, case["F_depths"], rtol=rtol, atol=1e-15)
        np.testing.assert_allclose(h.G_depths().values, case["G_depths"], rtol=rtol, atol=1e-15)
        assert list(F.index) == ["F%d" % i for i in range(len(case["F"]))]  # the caller's F is not mutated
    elif kind == "functional_homogeneity_mv":
        from statdepth_b200.homogeneity import FunctionalHomogeneity
        F = [pd.DataFrame(np.array(case["F"])[i]) for i in range(len(case["F"]))]
        G = [pd.DataFrame(np.array(case["G"])[i]) for i in range(len(case["G"]))]
        h = FunctionalHomogeneity(F, G, method=case["method"], containment="simplex", relax=case["relax"], quiet=True)
        got, want = float(h.homogeneity()), case["value"]
        assert (np.isnan(got) and np.isnan(want)) or np.isclose(got, want, rtol=rtol, atol=0)
        assert len(F) == len(case["F"])  # the caller's list is not mutated
    elif kind == "mahalanobis":
        df = pd.DataFrame(np.array(case["P"]))
        try:
            np.linalg.inv(np.cov(df, rowvar=True))
        except np.linalg.LinAlgError:  # this machine's LAPACK notices the singularity: so must the product
            import pytest
            with pytest.raises(np.linalg.LinAlgError):
                PointcloudDepth(df, containment="mahalanobis", to_compute=case.get("to_compute"))
            return
        res = PointcloudDepth(df, containment="mahalanobis", to_compute=case.get("to_compute"))
        # n == p (required, _pointcloud.py:160-161): the covariance of n points in n dimensions is singular, so the
        # "inverse" is whatever LAPACK makes of it (values ~1e16) and depends on the BLAS kernels of the machine.  The
        # check is therefore the reference's own formula (_pointcloud.py:163-172) recomputed HERE; the stored numbers
        # are compared only where they are the same machine's (same order of magnitude).
        mu, inv_cov = df.mean(), np.linalg.inv(np.cov(df, rowvar=True))
        idx = df.index if case.get("to_compute") is None else case["to_compute"]
        want = [np.dot((df.loc[p, :] - mu).T, np.dot(inv_cov, df.loc[p, :])) for p in idx]
        np.testing.assert_allclose(res.values, want, rtol=1e-9)
        assert len(res) == len(case["depths"])
        if np.allclose(want, case["depths"], rtol=1e-3):
            np.testing.assert_allclose(res.values, case["depths"], rtol=1e-9)
    else:
        raise AssertionError(kind)
