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
