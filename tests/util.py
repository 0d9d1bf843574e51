"""helpers shared by the parity tests"""
import numpy as np

TOL = {"f64": 1e-6, "f32": 1e-4}   # BASELINE.json north_star: 1e-6 relative (double Real), 1e-4 (float Real)
# The north star names influence matrix, source function and brightness.  The other three tracker
# outputs (tau_species_final, tau_absorber_final, species_col_dens) are sums of 4-point interpolants
# whose radial weight is (logf(r) - l0)/(l1 - l0): in float that quotient carries ~2e-4 of rounding noise
# (ulp(logf(3.5e8)) = 1.9e-6 over l1 - l0 ~ 1e-2), so two float builds that differ by one ulp in logf
# already disagree at the 1e-4 level on them.  They are held to 1e-3 in float, 1e-6 in double.
TOL_AUX = {"f64": 1e-6, "f32": 1e-3}


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|a|,|b|,floor) over entries where either is non-zero"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = (a != 0) | (b != 0)
    if not m.any():
        return 0.0
    den = np.maximum(np.maximum(np.abs(a[m]), np.abs(b[m])), floor)
    return float((np.abs(a - b)[m] / den).max())


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))


def assert_lists_equal(a, b):
    """(len, exits_bottom, entering, distance) tuples: integers and IEEE bits identical"""
    assert np.array_equal(a[0], b[0]), "boundary-list lengths differ"
    assert np.array_equal(a[1], b[1]), "exits_bottom flags differ"
    assert np.array_equal(a[2], b[2]), "entering voxel indices differ"
    assert same_bits(a[3], b[3]), "crossing distances differ in their IEEE bits"
