"""Size-independent properties of the path, checked on the oracle (CPU).  They restate the
debug-build asserts that are the reference's de-facto unit tests."""
import numpy as np

from util import rel_err


def test_row_sums_are_probabilities(synth, oraclebind):
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2)
    O = oraclebind.OracleModel(scn)
    O.build_rows()
    for e in range(2):
        K = O.K(e)
        assert (K >= 0).all()
        assert K.sum(axis=1).max() < 1.0       # emission_voxels.hpp:161 (commented assert)
        v = O.vectors(e)
        assert ((v["S0"] >= 0) & (v["S0"] <= 1)).all()     # holstein T is a probability, singlet_CFR.hpp:187-190


def test_zero_branching_gives_single_scattering(synth, oraclebind):
    """(I - w K) S = S0 with w = 0 must return S0 exactly"""
    scn = synth.make_scenario(8, 6, 4, 4, n_em=1)
    scn.em_scalars[0][0] = 0.0
    O = oraclebind.OracleModel(scn)
    O.build_rows()
    O.solve()
    v = O.vectors(0)
    assert np.array_equal(v["S"], v["S0"])


def test_optically_thin_limit(synth, oraclebind):
    """tau -> 0: S0 -> 1, K -> 0, S -> 1.  The brightness tends to g*N_col/1e9 * sqrt(T_ref/T)/sqrt(pi):
    the unnormalised T_int is clamped to the line-centre optical depth of the step
    (singlet_CFR.hpp:244-248), which is reference behaviour and is reproduced as is."""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    scn.vox_in[[0, 1, 4, 5]] *= 1e-10        # same geometry, densities scaled to tau << 1
    O = oraclebind.OracleModel(scn)
    O.build_rows()
    O.solve()
    v = O.vectors(0)
    lit = v["S0"] > 0
    assert np.abs(v["S0"][lit] - 1).max() < 1e-4
    assert np.abs(v["S"][lit] - 1).max() < 1e-4
    assert O.K(0).max() < 1e-6
    locs, dirs = synth.fake_image(30 * synth.rMars, 30, 12)
    _, b = O.brightness(locs, dirs, 10)
    hit = (b[0, 3] > 0) & (b[0, 2] >= 0)
    g = scn.em_scalars[0][3]
    thin = g * b[0, 3][hit] / 1e9
    ratio = b[0, 0][hit] / thin
    assert (ratio < 1 / np.sqrt(np.pi) * 1.3).all()          # T <= T_ref*1.6 everywhere in this atmosphere
    clear = ratio > 0.5                                       # lines of sight clear of the planet's shadow
    assert clear.any()
    assert np.abs(ratio[clear] * np.sqrt(np.pi) - 1).max() < 0.3


def test_linearity_in_g_factor(synth, oraclebind):
    scn = synth.make_scenario(8, 6, 4, 4, n_em=1)
    O = oraclebind.OracleModel(scn)
    O.build_rows()
    O.solve()
    locs, dirs = synth.random_los(200, seed=3)
    _, b1 = O.brightness(locs, dirs, 10)
    scn2 = synth.make_scenario(8, 6, 4, 4, n_em=1)
    scn2.em_scalars[0][3] *= 2.0
    O2 = oraclebind.OracleModel(scn2)
    O2.build_rows()
    O2.solve()
    _, b2 = O2.brightness(locs, dirs, 10)
    assert rel_err(2.0 * b1[0, 0], b2[0, 0]) < 1e-13
    assert np.array_equal(b1[0, 1:], b2[0, 1:])


def test_boundary_list_continuity(synth, oraclebind):
    """boundary_set::check (boundaries.hpp:235-270): consecutive voxels differ by one step in
    exactly one dimension; rays leave through the top or the bottom"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    O = oraclebind.OracleModel(scn)
    ln, eb, ent, dist = O.traverse_voxel_rays()
    nsb1 = scn.n_sb - 1
    pos = 0
    for n in ln:
        e = ent[pos:pos + n]
        d = dist[pos:pos + n]
        assert n >= 2 and e[-1] == -1 and (e[:-1] >= 0).all()
        assert (np.diff(d) >= 0).all() and d[0] == 0.0
        ri, si = e[:-1] // nsb1, e[:-1] % nsb1
        step = np.abs(np.diff(ri)) + np.abs(np.diff(si))
        assert (step == 1).all()
        pos += n
