"""Golden vectors produced by the reference's own source (tests/golden/make_golden.py):
the oracle (CPU, always) and the CUDA path (-m gpu) must reproduce them."""
import numpy as np
import pytest

from golden_util import GOLDEN, load_golden
from util import TOL, UNDERFLOW, TOL_AUX, assert_lists_equal, rel_err, same_bits


def check_against_golden(M, scn, z, prec, exact):
    """exact=True (oracle): bit-identical; exact=False (CUDA): indices/distances bit-identical,
    floating-point results within the north-star tolerance."""
    tol = TOL[prec]
    g = M.grid()
    for k in ("sza_boundaries", "pts_radii", "pts_sza", "ray_theta", "ray_phi", "ray_domega"):
        assert same_bits(g[k], z["grid_" + k]), k
    assert_lists_equal(M.traverse_voxel_rays(), (z["vr_len"], z["vr_eb"], z["vr_ent"], z["vr_dist"]))
    _, nsteps = M.build_rows()
    assert nsteps == int(z["n_steps"])
    for e in range(2):
        K = M.K(e)
        if exact:
            assert same_bits(K, z[f"K{e}"])
        else:
            # entries at the underflow edge of Real (< 1e-290 / 1e-30, against row sums of order 1) are compared as
            # zeros: the device's exp does not walk through the denormals (fastmath.cuh), same rule as test_gpu_parity
            floor = 1e-290 if prec == "f64" else 1e-30
            Kz = z[f"K{e}"]
            assert np.array_equal(np.abs(K) > floor, np.abs(Kz) > floor)
            assert rel_err(np.where(np.abs(K) > floor, K, 0.0), np.where(np.abs(Kz) > floor, Kz, 0.0)) < tol
        v = M.vectors(e, want_S=False) if not exact else M.vectors(e)
        for k in ("S0", "tau_species_ss", "tau_absorber_ss"):
            if exact:
                assert same_bits(v[k], z[f"vec{e}_{k}"]), k
            else:
                assert rel_err(v[k], z[f"vec{e}_{k}"], floor=UNDERFLOW[prec]) < tol, k
    M.solve()
    for e in range(2):
        Sg = z[f"vec{e}_S"]
        floor = 1e-30 if prec == "f64" else float(np.abs(Sg).max())
        assert rel_err(M.vectors(e)["S"], Sg, floor=floor) < (1e-7 if prec == "f64" else 1e-3 if exact else tol)
        M.set_sourcefn(e, Sg)
    locs, dirs = z["los_loc"], z["los_dir"]
    lists = M.traverse_los(locs, dirs)
    assert_lists_equal(lists[:4], (z["los_len"], z["los_eb"], z["los_ent"], z["los_dist"]))
    for nsub in (10, 0):
        _, b = M.brightness(locs, dirs, nsub)
        ref = z[f"brightness_nsub{nsub}"]
        if exact:
            assert same_bits(b, ref)
        else:
            assert np.array_equal(b[:, 2] == -1, ref[:, 2] == -1)      # lines of sight that hit the planet
            for q in range(4):
                assert rel_err(b[:, q], ref[:, q], floor=1e-300) < (tol if q == 0 else TOL_AUX[prec]), (nsub, q)


@pytest.mark.parametrize("name,prec", GOLDEN)
def test_oracle_reproduces_golden(synth, oraclebind, name, prec):
    scn, z = load_golden(synth, name, prec)
    check_against_golden(oraclebind.OracleModel(scn, prec), scn, z, prec, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,prec", GOLDEN)
def test_cuda_reproduces_golden(synth, binding, name, prec):
    scn, z = load_golden(synth, name, prec)
    check_against_golden(binding.GpuModel(scn, prec), scn, z, prec, exact=False)
