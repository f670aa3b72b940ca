"""Engine-wide knobs that have no counterpart in the reference's signatures."""

# Absolute tolerance band of the closed-simplex test.  The reference decides containment with
# scipy.optimize.linprog (statdepth/depth/calculations/_containment.py:171), whose primal
# feasibility tolerance (HiGHS, scipy 1.18) accepts points up to ~1e-7 outside a face.
_SIMPLEX_TOL = 1e-7


def get_simplex_tolerance() -> float:
    return _SIMPLEX_TOL


def set_simplex_tolerance(tol: float) -> None:
    """0.0 selects exact closed-simplex sign tests."""
    global _SIMPLEX_TOL
    tol = float(tol)
    if not tol >= 0.0:
        raise ValueError("tolerance must be >= 0")
    _SIMPLEX_TOL = tol


# Largest enumeration (subsets x queries [x time points]) the engine will start.  Simplicial depth in 2-D and
# relaxed multivariate simplex depth in 2-D are COUNTED (O(n log n) per query) and never hit this limit; d = 3,
# strict multivariate depth and Oja depth enumerate C(n-1, k) subsets per query like the reference does.
_MAX_ENUMERATION = 5e13


def get_max_enumeration() -> float:
    return _MAX_ENUMERATION


def set_max_enumeration(limit: float) -> None:
    global _MAX_ENUMERATION
    _MAX_ENUMERATION = float(limit)


def check_enumeration(work: float, what: str) -> None:
    if work > _MAX_ENUMERATION:
        raise NotImplementedError(
            '%s would enumerate %.3g simplices; the B200 engine enumerates like the reference for this case and '
            'refuses more than %.3g (statdepth_b200.settings.set_max_enumeration raises the limit)'
            % (what, work, _MAX_ENUMERATION))
