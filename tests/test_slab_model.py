"""CPU: a numpy model of the slab rank path's arithmetic (csrc/mbd_slab.cuh) -- the sample-built monotone code, the bins,
the exactness argument.  The kernels themselves are checked on the GPU (tests/test_gpu_parity.py::test_mbd_slab_path);
this model pins the INVARIANTS they rely on, with the same integer formulas:
  * code(x) is non-decreasing in x for every input (also outside the sampled range),
  * "lower bin" / "lower 14-bit key in the same bin" therefore imply a strictly smaller value, so ranks rebuilt from
    (bin start + position among the bin's keys, exact values only where keys are equal) are the true strict ranks,
  * on continuous data the validated capacities hold with margin: a CTA's share of the row, bins of at most 255
    values, the pair work of the bins of more than 16."""
import numpy as np
import pytest

SL_BUCKETS, SL_SHIFT, SL_SAMPLE, SL_TRIM, CHUNK = 256, 17, 16384, 4, 256


def plan(n):
    from statdepth_b200 import build
    from statdepth_b200._engine import mbd_plan
    build.build()  # no-op when libsdepth.so is up to date
    p = mbd_plan(n)
    assert p["slab"] == 1
    return p["ctas_per_row"], p["bins_per_cta"], p["entries_per_cta"]


def table_and_codes(x):
    """What mbd_slab_table_kernel + slab_code compute for one row (x*s + c in two roundings instead of one fma:
    both are monotone, which is all the argument needs)."""
    n = len(x)
    G, nbc, ecap = plan(n)
    NB = G * nbc
    gap = (n - CHUNK) // (SL_SAMPLE // CHUNK - 1)
    starts = np.arange(SL_SAMPLE // CHUNK) * gap
    sample = np.concatenate([x[s:s + CHUNK] for s in starts])
    e = np.arange(1024)
    x0 = np.median([x[0], x[n >> 1], x[n - 1]])              # row_reference: the range sample is sorted as float(x - x0)
    sorted1024 = x0 + np.sort((x[(e >> 4) * gap + ((e & 15) << 4)] - x0).astype(np.float32)).astype(np.float64)
    qlo, qhi = sorted1024[SL_TRIM], sorted1024[1023 - SL_TRIM]
    span = qhi - qlo
    lo, hi = qlo - 0.35 * span, qhi + 0.35 * span
    s = (SL_BUCKETS - 2) * 4294967296.0 / (hi - lo)
    c = (4503599627370496.0 + 4294967296.0) - lo * s

    def split(v):
        d = v * s + c
        bits = np.ascontiguousarray(d).view(np.uint64)
        top = (bits >> np.uint64(32)).astype(np.int64)
        top = np.where(top >= 2 ** 31, top - 2 ** 32, top)            # the high word as a signed int
        k = np.clip(top, 0x43300000, 0x43300000 + SL_BUCKETS - 1) - 0x43300000
        return k, (bits & np.uint64(0xffffffff)).astype(np.uint64)

    ks, _ = split(sample)
    counts = np.bincount(ks, minlength=SL_BUCKETS)
    excl = np.concatenate([[0], np.cumsum(counts)[:-1]])
    top = (NB << SL_SHIFT) - 1
    scl = float(top) / float(SL_SAMPLE)
    c0 = np.minimum((excl * scl).astype(np.uint64), top)
    c1 = np.minimum(((excl + counts) * scl).astype(np.uint64), top)
    C, D = c0, c1 - c0
    D[0] = D[-1] = 0
    k, frac = split(x)
    code = C[k] + ((frac * D[k]) >> np.uint64(32))
    assert int(code.max()) < (NB << SL_SHIFT)
    return code.astype(np.int64), (G, nbc, ecap)


def ranks_from_codes(x, code):
    """b = #values in lower bins + #lower keys in the bin + #smaller values among equal keys; a likewise."""
    n = len(x)
    bink = code >> 3          # bin (code >> 17) and the 14-bit key ((code >> 3) & 0x3fff) in one integer
    order = np.lexsort((x, bink))
    bs, xs = bink[order], x[order]
    first_of_key = np.searchsorted(bs, bs, side="left")      # entries with a lower (bin, key)
    last_of_key = np.searchsorted(bs, bs, side="right")
    below = np.empty(n, dtype=np.int64)
    above = np.empty(n, dtype=np.int64)
    lt = np.array([np.searchsorted(xs[f:l], v, side="left") for f, l, v in zip(first_of_key, last_of_key, xs)])
    le = np.array([np.searchsorted(xs[f:l], v, side="right") for f, l, v in zip(first_of_key, last_of_key, xs)])
    below[order] = first_of_key + lt
    above[order] = n - first_of_key - le
    return below, above


@pytest.mark.parametrize("kind", ["normal", "walk", "expo", "t3", "bimodal", "shifted", "outliers", "sparse_ties"])
def test_code_is_monotone_and_ranks_are_exact(kind):
    rng = np.random.default_rng(len(kind))
    n = 40_000
    x = {"normal": lambda: rng.standard_normal(n), "walk": lambda: rng.standard_normal((30, n)).cumsum(0)[-1],
         "expo": lambda: rng.standard_exponential(n), "t3": lambda: rng.standard_t(3, n),
         "bimodal": lambda: np.concatenate([rng.standard_normal(n // 2), 5 + 0.1 * rng.standard_normal(n - n // 2)]),
         "shifted": lambda: 1e6 + 1e-3 * rng.standard_normal(n),
         "outliers": lambda: np.where(np.arange(n) % 997 == 0, 1e9, 1.0) * rng.standard_normal(n),
         "sparse_ties": lambda: np.repeat(rng.standard_normal(n // 2), 2)[rng.permutation(n)]}[kind]()
    code, _ = table_and_codes(x)
    order = np.argsort(x, kind="stable")
    assert (np.diff(code[order]) >= 0).all()                  # monotone, also for the clamped values
    below, above = ranks_from_codes(x, code)
    xs = np.sort(x)
    assert (below == np.searchsorted(xs, x, side="left")).all()
    assert (above == n - np.searchsorted(xs, x, side="right")).all()


@pytest.mark.parametrize("n", [16384, 50_000, 100_000, 131_072])
def test_validated_capacities_hold_on_continuous_rows(n):
    rng = np.random.default_rng(n)
    for x in (rng.standard_normal(n), rng.standard_normal((40, n)).cumsum(0)[-1], rng.random(n),
              rng.standard_exponential(n)):
        code, (G, nbc, ecap) = table_and_codes(x)
        bins = np.bincount(code >> SL_SHIFT, minlength=G * nbc)
        share = bins.reshape(G, nbc).sum(1)
        big = bins[bins > 16].astype(np.int64)
        assert share.max() <= ecap and bins.max() <= 255 and int((big * big).sum()) <= 1 << 18
        assert len(big) <= 0.02 * G * nbc                      # bins the warp path ranks: a percent or so
        assert share.max() <= 1.04 * n / G                     # sample-built table: shares within a few percent
