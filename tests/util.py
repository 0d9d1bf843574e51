"""helpers shared by the parity tests"""
import numpy as np

TOL = {"f64": 1e-6, "f32": 1e-4}   # BASELINE.json north_star: 1e-6 relative (double Real), 1e-4 (float Real)
# The north star names influence matrix, source function and brightness.  The other three tracker
# outputs (tau_species_final, tau_absorber_final, species_col_dens) are sums of 4-point interpolants
# whose radial weight is (log(r) - l0)/(l1 - l0): in float one ulp of log(r) moves that weight by ~3e-5,
# so the device rounds the DOUBLE log to float (what the host libm's logf returns in all but rare
# cases); measured worst case over the test sets is 5.3e-5.  They are held to 2e-4 in float.
TOL_AUX = {"f64": 1e-6, "f32": 2e-4}


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


# Results below this magnitude are compared as zeros: quantities of order 1 (transmissions, probabilities) that have
# underflowed through ~700 e-foldings.  The host libm walks through the denormals to exactly 0; the device's exp
# (fastmath.cuh) bottoms out at ~1e-301 instead.
UNDERFLOW = {"f64": 1e-290, "f32": 1e-30}


def write_iph_table_file(tab, path):
    """The reference's Quemerais table file layout (ipbackgroundCFR_fun.f:107-164) written from
    in-memory tables: `KMAX LMAX INF`, the 8-number header (V0 VLON VDEC TEMP AMU TDUR AN DINF), then four angle blocks
    (5, 5, 5, 4 columns) for each of DANS, SOT, SO(:,:,1), SN(:,:,1) -- each block one ANG row and KMAX rows `ALT v...` --
    and for every further density at infinity a header and the blocks of SO / SN.  Values are printed with 9 significant
    digits, which round-trips float32.  Lets the parsers be exercised where the reference's own file is absent."""
    kmax, lmax, ninf = int(tab["kmax"]), int(tab["lmax"]), int(tab["ninf"])
    alt, ang = np.asarray(tab["alt_au"], np.float32), np.asarray(tab["ang"], np.float32)
    cols = [(0, 5), (5, 10), (10, 15), (15, 19)]
    f9 = lambda v: "%.9g" % float(v)

    def header(ii):
        return " ".join(f9(v) for v in (20.0, 254.0, 7.5, tab["temp"], 0.99, 1.2e6, 0.0, tab["dinf_cm3"][ii])) + "\n"

    def blocks(a):
        out = []
        for c0, c1 in cols:
            out.append("      " + " ".join(f9(v) for v in ang[c0:c1]) + "\n")
            for k in range(kmax):
                out.append(f9(alt[k]) + " " + " ".join(f9(v) for v in a[k, c0:c1]) + "\n")
        return "".join(out)

    with open(path, "w") as f:
        f.write(f"{kmax} {lmax} {ninf}\n")
        f.write(header(0))
        for a in (tab["dans"], tab["sot"], tab["so"][0], tab["sn"][0]):
            f.write(blocks(np.asarray(a, np.float32)))
        for ii in range(1, ninf):
            f.write(header(ii))
            f.write(blocks(np.asarray(tab["so"][ii], np.float32)))
            f.write(blocks(np.asarray(tab["sn"][ii], np.float32)))
