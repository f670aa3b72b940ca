"""CPU: the Python host layer (validation, labels, K-sampling replay, float assembly, result types)
against the reference's golden vectors, with the compute answered by the oracle-backed TEST DOUBLE
(tests/conftest.py::OracleBackedEngine).  The product itself never routes through the oracle."""
import numpy as np
import pandas as pd
import pytest

from api_cases import check_case, check_next_case, load_next_cases
from statdepth_b200 import DepthDegeneracy, EngineUnavailable, FunctionalDepth, PointcloudDepth
from statdepth_b200.homogeneity import FunctionalHomogeneity, P1_homogeneity, P2_homogeneity
from statdepth_b200.testing import (generate_noisy_multivariate, generate_noisy_pointcloud,
                                    generate_noisy_univariate)


def test_no_cpu_fallback():
    """Without a GPU the product refuses to compute (it must not silently fall back)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(EngineUnavailable):
        FunctionalDepth([generate_noisy_univariate(seed=0)])


def test_all_golden_cases(host_on_oracle, golden):
    n = 0
    for case in golden.values():
        if case["kind"] in ("functional", "multivariate", "pointcloud"):
            check_case(case, FunctionalDepth, PointcloudDepth)
            n += 1
    assert n >= 30


def test_result_types_like_reference_tests(host_on_oracle):
    """The reference's own tests (tests/test_statdepth.py:22-73) assert exactly these types."""
    df = generate_noisy_univariate(seed=1)
    bd = FunctionalDepth([df], containment='r2')
    for obj in (bd, bd.ordered(), bd.median(), bd.deepest(n=2), bd.outlying(n=2)):
        assert isinstance(obj, pd.Series)
    assert isinstance(FunctionalDepth([df], K=5, containment='r2'), pd.Series)
    pc = generate_noisy_pointcloud(n=10, d=2, seed=2)
    for c in ('l1', 'simplex', 'oja'):
        r = PointcloudDepth(pc, containment=c)
        for obj in (r, r.ordered(), r.median(), r.deepest(n=2), r.outlying(n=2)):
            assert isinstance(obj, pd.Series)
        assert isinstance(PointcloudDepth(pc, K=2, containment=c), pd.Series)
    mv = FunctionalDepth(generate_noisy_multivariate(seed=3), containment='simplex')
    assert isinstance(mv, pd.Series) and isinstance(mv.ordered(), pd.Series)
    assert list(mv.values) == [0.0, 1.0, 1.0, 1.0, 0.0]  # reference output for seed 3
    assert bd.get_deepest_data().shape == (20, 1)
    assert bd.drop_outlying_data(n=2).shape == (20, 18)
    assert len(bd.quartile()) == 10


def test_k_sampled_matches_reference_stream(host_on_oracle, golden):
    """Same np.random.seed -> same blocks -> same sampled depths as the reference (_functional.py:154-186)."""
    case = golden["walk_K3_seed123"]
    df = pd.DataFrame(np.array(case["X"]))
    np.random.seed(case["np_seed"])
    res = FunctionalDepth([df], K=case["K"], **case["kwargs"])
    assert [int(i) for i in res.index] == case["index"]
    np.testing.assert_allclose(res.values, case["depths"], rtol=1e-12)


def test_homogeneity_matches_reference(host_on_oracle, golden):
    for m in ("p1", "p2", "p3"):
        case = golden["functional_%s" % m]
        F = pd.DataFrame(np.array(case["F"]), columns=["F%d" % i for i in range(7)])
        G = pd.DataFrame(np.array(case["G"]), columns=["G%d" % i for i in range(6)])
        keep = F.copy()
        h = FunctionalHomogeneity([F], [G], method=m, quiet=True).homogeneity()
        np.testing.assert_allclose(float(np.asarray(h).ravel()[0]), case["value"], rtol=1e-12)
        assert F.equals(keep)  # documented difference: the caller's F is not mutated
    assert isinstance(float(P1_homogeneity(F, G)), float)
    assert float(P2_homogeneity(F, G)) >= 0.0


def test_validation_errors():
    """_handle_depth_errors (_helper.py:34-107): same exception types, raised before any GPU work."""
    df = generate_noisy_univariate(seed=0)
    with pytest.raises(ValueError, match='data must be passed as a list'):
        FunctionalDepth(df)
    with pytest.raises(ValueError, match='J must be an integer'):
        FunctionalDepth([df], J=2.0)
    with pytest.raises(ValueError, match='greater than or equal to 2'):
        FunctionalDepth([df], J=1)
    with pytest.raises(ValueError, match='relax must be of type bool'):
        FunctionalDepth([df], relax=1)
    with pytest.raises(ValueError, match='No data passed'):
        FunctionalDepth([])
    with pytest.raises(ValueError, match='less than the number of observations'):
        FunctionalDepth([df], J=20)  # J compared with the number of ROWS (reference quirk)
    mv = generate_noisy_multivariate(seed=0)
    with pytest.raises(ValueError, match="'r2' is invalid for multivariate"):
        FunctionalDepth(mv, containment='r2')
    with pytest.raises(DepthDegeneracy):
        FunctionalDepth(mv[:4], containment='simplex')  # needs d + 2 = 5 curves
    with pytest.raises(TypeError):
        FunctionalDepth([df], containment='simplex')    # reference: TypeError escapes from _helper.py:92-93
    with pytest.raises(ValueError, match='is invalid'):
        FunctionalDepth([df], containment='nope')
    with pytest.raises(ValueError, match='incorrect number of parameters'):
        FunctionalDepth([df], containment=lambda a, b: 0.0)
    with pytest.raises(NotImplementedError):
        FunctionalDepth([df], containment=lambda data, curve, relax: 0.0)
    with pytest.raises(NotImplementedError):
        FunctionalDepth([df], containment='r2_enum')
    with pytest.raises(ValueError, match='numeric dtypes'):
        FunctionalDepth([df.astype(object).replace(df.iloc[0, 0], 'x')], deep_check=True)
    with pytest.raises(DepthDegeneracy, match='Block size'):
        FunctionalDepth([df], K=50)
    with pytest.raises(ValueError, match='not a valid containment'):
        PointcloudDepth(generate_noisy_pointcloud(seed=0), containment='nope')


def test_generators_replay_reference_stream(golden):
    case = golden["generator_default_seed4"]
    assert np.array_equal(generate_noisy_univariate(seed=4).values, np.array(case["X"]))
    assert np.array_equal(generate_noisy_pointcloud(n=10, d=2, seed=2).values,
                          np.array(golden["simplex_gen_seed2"]["P"]))
    F = np.stack([d.values for d in generate_noisy_multivariate(seed=0)])
    assert np.array_equal(F, np.array(golden["generator_deg_seed0"]["F"]))


def test_j4_relaxed_from_ranks(host_on_oracle, oracle):
    from math import comb
    X = np.random.default_rng(3).standard_normal((9, 11)).cumsum(0)
    res = FunctionalDepth([pd.DataFrame(X)], J=4, relax=True)
    _, b, a = oracle.mbd_counts_all(X, want_ranks=True)
    exp = np.zeros(11)
    for j in (2, 3, 4):
        s = np.array([sum(comb(10, j) - comb(int(bb), j) - comb(int(aa), j) for bb, aa in zip(b[:, c], a[:, c]))
                      for c in range(11)], dtype=np.float64)
        exp = exp + s / 9 / comb(11, j)
    np.testing.assert_allclose(res.values, exp, rtol=1e-13)
    with pytest.raises(NotImplementedError):
        FunctionalDepth([pd.DataFrame(X)], J=4, relax=False)


def test_permutation_test_batched_equals_loop(host_on_oracle, monkeypatch):
    """permutation_test: the batched evaluation (2-3 engine calls for all permutations) reproduces the
    per-permutation FunctionalHomogeneity loop exactly, including tie-breaking of the deepest curve."""
    import statdepth_b200._engine as eng_mod
    from statdepth_b200.homogeneity import permutation_test
    monkeypatch.setattr(eng_mod, "get_engine", lambda device=None: host_on_oracle)
    rng = np.random.default_rng(5)
    F = pd.DataFrame(rng.standard_normal((12, 9)).cumsum(0))
    G = pd.DataFrame(np.round(rng.standard_normal((12, 8)).cumsum(0)) + 1.0)  # ties included
    for method in ("p1", "p2"):
        for relax in (True, False):
            a = permutation_test(F, G, method=method, B=12, seed=3, relax=relax, batched=True)
            b = permutation_test(F, G, method=method, B=12, seed=3, relax=relax, batched=False)
            assert a["observed"] == b["observed"] and a["p_value"] == b["p_value"]
            assert a["null"].tolist() == b["null"].tolist()


def test_enumeration_guard(host_on_oracle):
    """Cases the engine can only enumerate (d = 3, strict multivariate, Oja) refuse oversized jobs up front;
    2-D simplicial depth and relaxed 2-D multivariate depth are counted and have no such limit."""
    from statdepth_b200 import settings
    rng = np.random.default_rng(0)
    old = settings.get_max_enumeration()
    try:
        settings.set_max_enumeration(1e4)
        with pytest.raises(NotImplementedError, match='enumerate'):
            PointcloudDepth(pd.DataFrame(rng.standard_normal((30, 3))), containment='simplex')
        with pytest.raises(NotImplementedError, match='enumerate'):
            PointcloudDepth(pd.DataFrame(rng.standard_normal((60, 2))), containment='oja')
        curves = [pd.DataFrame(rng.standard_normal((4, 2))) for _ in range(40)]
        with pytest.raises(NotImplementedError, match='enumerate'):
            FunctionalDepth(curves, containment='simplex', relax=False)
        assert len(PointcloudDepth(pd.DataFrame(rng.standard_normal((12, 2))), containment='simplex')) == 12
    finally:
        settings.set_max_enumeration(old)


def test_next_rows_against_the_reference(host_on_oracle):
    """K-sampled point-cloud depth (blocks replayed from the global RNG, one batched call), point-cloud
    homogeneity p1..p3 and Mahalanobis depth: outputs of the unmodified reference (make_golden_next.py)."""
    from statdepth_b200.homogeneity import PointcloudHomogeneity
    cases = load_next_cases()
    assert len(cases) >= 12
    for case in cases:
        check_next_case(case, PointcloudDepth, PointcloudHomogeneity)


def test_functional_p3_batched_equals_loop(host_on_oracle, golden):
    """p3 evaluates its |G| single-query depth runs as one batched call; K forces the reference's loop."""
    case = golden["functional_p3"]
    F = pd.DataFrame(np.array(case["F"]), columns=["F%d" % i for i in range(7)])
    G = pd.DataFrame(np.array(case["G"]), columns=["G%d" % i for i in range(6)])
    h = FunctionalHomogeneity([F], [G], method="p3", quiet=True).homogeneity()
    np.testing.assert_allclose(float(np.asarray(h).ravel()[0]), case["value"], rtol=1e-12)
    for relax in (True, False):
        from statdepth_b200 import homogeneity as H
        batched = H._depths_of_each_in(F, G, 3, relax)
        loop = []
        for col in G.columns:
            Fc = F.copy()
            Fc.loc[:, col] = G.loc[:, col].values
            loop.append(FunctionalDepth([Fc], to_compute=[col], J=3, relax=relax).loc[col])
        np.testing.assert_allclose(batched, loop, rtol=1e-13)


def test_mbd_plan_geometry(monkeypatch):
    """sd_mbd_plan (host arithmetic of the library, no GPU): the slab path's geometry stays inside what its kernels
    assume for every eligible row length -- shared memory within the 227 KB opt-in limit, 16-bit bin starts, bins per
    CTA a multiple of 1024 with 5 .. 9.5 values per bin -- and everything else goes to the part pipeline."""
    from statdepth_b200 import build
    from statdepth_b200._engine import mbd_plan
    build.build()  # no-op when libsdepth.so is up to date
    for k in ("SD_MBD_PATH", "SD_MBD_SLAB_MIN", "SD_MBD_SLAB_G", "SD_MBD_SLAB_THREADS"):
        monkeypatch.delenv(k, raising=False)
    for n in (1, 200, 16382, 131074, 250_000):
        assert mbd_plan(n)["slab"] == 0
    assert mbd_plan(20_001, 20_001)["slab"] == 0 and mbd_plan(20_000, 20_001)["slab"] == 0  # odd leading dimension
    rng = __import__("numpy").random.default_rng(0)
    sizes = [16384, 16386, 20_000, 50_000, 65_536, 100_000, 131_070, 131_072] + \
        [int(2 * v) for v in rng.integers(8192, 65536, size=200)]
    for n in sizes:
        p = mbd_plan(n)
        assert p["slab"] == 1, n
        G, nbc, ecap = p["ctas_per_row"], p["bins_per_cta"], p["entries_per_cta"]
        assert 1 <= G <= 8 and nbc % 1024 == 0 and nbc >= 1024
        assert p["smem_rank"] == 4 * (ecap + 16) + 4 * nbc and p["smem_rank"] <= 232448 - 256
        assert p["smem_hist"] == 4 * G * nbc and p["smem_hist"] + 2192 <= 232448 // 2
        assert ecap <= 65535 and ecap * G >= n * 1.05        # a CTA's share of the row with slack, 16-bit starts
        assert G * nbc * (1 << 17) < 2 ** 32                   # codes: bin << 17 | fraction in 32 bits
        assert 4.0 <= n / (G * nbc) <= 9.5, (n, G, nbc)
    assert mbd_plan(100_000) == {"slab": 1, "ctas_per_row": 3, "bins_per_cta": 4096, "entries_per_cta": 35545,
                                 "smem_rank": 4 * (35545 + 16) + 4 * 4096, "smem_hist": 4 * 3 * 4096}
    monkeypatch.setenv("SD_MBD_PATH", "parts")
    assert mbd_plan(100_000)["slab"] == 0
