"""CUDA path (through the C ABI, libb200rt.so) against the oracle on the same seeded inputs.

Bar (BASELINE.json north_star): voxel traversal and intersection indices bit-exact;
influence matrix, source function and brightness within 1e-6 relative (double Real) and
1e-4 (float Real).  Run on the B200 box: python -m pytest tests -m gpu
"""
import os

import numpy as np
import pytest

from util import TOL, UNDERFLOW, TOL_AUX, assert_lists_equal, rel_err

pytestmark = pytest.mark.gpu


def compare_models(synth, O, G, scn, prec, los_sets):
    tol = TOL[prec]
    assert_lists_equal(O.traverse_voxel_rays(), G.traverse_voxel_rays())
    _, ns_o = O.build_rows()
    _, ns_g = G.build_rows()
    assert ns_o == ns_g
    for e in range(scn.n_em):
        Ko, Kg = O.K(e), G.K(e)
        # zero pattern: entries at the underflow edge of Real (< 1e-290 / 1e-30, against row sums of order 1)
        # depend on the order in which the products of a step are formed (the device uses per-voxel
        # kappa / wratio tables, influence.cu) and are compared as zeros
        floor = 1e-290 if prec == "f64" else 1e-30
        assert np.array_equal(np.abs(Ko) > floor, np.abs(Kg) > floor)
        Ko, Kg = np.where(np.abs(Ko) > floor, Ko, 0.0), np.where(np.abs(Kg) > floor, Kg, 0.0)
        assert rel_err(Ko, Kg) < tol
        vo, vg = O.vectors(e), G.vectors(e, want_S=False)
        for k in ("S0", "tau_species_ss", "tau_absorber_ss"):
            assert rel_err(vo[k], vg[k], floor=UNDERFLOW[prec]) < tol, k
        assert np.array_equal(vo["S0"] > UNDERFLOW[prec], vg["S0"] > UNDERFLOW[prec])   # shadowed voxels
    O.solve()
    res = G.solve()
    for e in range(scn.n_em):
        assert res[e] < 1e-12                                          # FP64 solve in both precisions
        So = O.vectors(e)["S"]
        floor = 1e-30 if prec == "f64" else float(np.abs(So).max())
        assert rel_err(So, G.vectors(e)["S"], floor=floor) < tol
        G.set_sourcefn(e, So)
    for locs, dirs in los_sets:
        a, b = O.traverse_los(locs, dirs), G.traverse_los(locs, dirs)
        assert_lists_equal(a[:4], b[:4])
        for nsub in (10, 0, 4):
            _, bo = O.brightness(locs, dirs, nsub)
            _, bg = G.brightness(locs, dirs, nsub)
            assert np.array_equal(bo[:, 2] == -1, bg[:, 2] == -1)
            for q in range(4):
                assert rel_err(bo[:, q], bg[:, q], floor=1e-300) < (tol if q == 0 else TOL_AUX[prec]), (nsub, q)


@pytest.mark.parametrize("prec", ["f64", "f32"])
@pytest.mark.parametrize("shape", [(8, 6, 4, 4), (12, 8, 5, 6), (20, 12, 6, 8)])
def test_small_grids(synth, binding, oraclebind, prec, shape):
    scn = synth.make_scenario(*shape, n_em=2, sza_T_contrast=0.1)
    O = oraclebind.OracleModel(scn, prec)
    G = binding.GpuModel(scn, prec)
    compare_models(synth, O, G, scn, prec, [synth.fake_image(30 * synth.rMars, 30, 24), synth.random_los(800)])


@pytest.mark.parametrize("szamethod,raymethod", [(0, 0), (0, 1), (1, 0)])
def test_other_grid_methods(synth, binding, oraclebind, szamethod, raymethod):
    """szamethod_uniform and Gauss-Legendre ray angles (grid_spherical...hpp:314-319,369-372)"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1, szamethod=szamethod, raymethod=raymethod)
    O = oraclebind.OracleModel(scn, "f64")
    G = binding.GpuModel(scn, "f64")
    compare_models(synth, O, G, scn, "f64", [synth.random_los(300)])


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_default_grid_D(synth, binding, oraclebind, prec):
    """BASELINE.json configs[0]: default 40x20x7x12 grid, Ly alpha + Ly beta"""
    scn = synth.make_scenario(40, 20, 7, 12, n_em=2)
    O = oraclebind.OracleModel(scn, prec)
    G = binding.GpuModel(scn, prec)
    compare_models(synth, O, G, scn, prec, [synth.fake_image(30 * synth.rMars, 30, 40), synth.random_los(3000)])


def test_row_sharding_equals_full(synth, binding):
    """rows built in ragged pieces (what each rank does under --gpus N) equal one full pass"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    Kfull = [G.K(e).copy() for e in range(2)]
    G2 = binding.GpuModel(scn, "f64")
    n = G2.n_vox
    total = 0
    for a, b in ((0, 1), (1, 30), (30, 30), (30, n)):     # includes an empty range
        _, ns = G2.build_rows(a, b)
        total += ns
    _, ns_full = G.build_rows()
    assert total == ns_full
    for e in range(2):
        assert rel_err(Kfull[e], G2.K(e)) < 1e-13         # fp64 RED ordering only


def test_interleaved_shards_equal_full(synth, binding):
    """b200rt_influence_ranges: the interleaved source-voxel shards of a multi-GPU build (one launch set per batch
    through the slot -> voxel map), including empty shards and a single-range call, equal one full pass"""
    import importlib
    multi = importlib.import_module("3d_planetary_rt_model_b200.multi")
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2, sza_T_contrast=0.05)
    G = binding.GpuModel(scn, "f64")
    _, ns_full = G.build_rows()
    Kfull = [G.K(e).copy() for e in range(2)]
    G2 = binding.GpuModel(scn, "f64")
    n, total = G2.n_vox, 0
    for rank in range(3):
        G2.ctx.influence_ranges(multi.partition_interleaved(n, 3, rank, 5))
        total += G2.ctx.last_step_count()
    assert total == ns_full
    for e in range(2):
        assert rel_err(Kfull[e], G2.K(e)) < 1e-13
    G2.ctx.influence_ranges([(0, n)])
    assert G2.ctx.last_step_count() == ns_full
    G2.ctx.influence_ranges([])                            # nothing to do: only the single-scattering rays run
    assert G2.ctx.last_step_count() == 0
    with pytest.raises(binding.B200RTError):
        G2.ctx.influence_ranges([(5, 9), (7, 12)])         # overlapping ranges are refused


def test_pipelined_host_brightness_equals_resident(synth, binding):
    """b200rt_brightness with host arrays runs as a pipeline of batches (upload / kernels / download on three streams,
    b200rt_api.cu brightness_impl); it must return exactly what upload -> brightness_resident -> download returns,
    for a line-of-sight count that spans several batches with a ragged last one, and with a null output skipped"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2)
    G = binding.GpuModel(scn, "f64")
    for e in range(2):
        G.set_sourcefn(e, np.linspace(1.0, 0.05, scn.n_vox) * (1 + e))
    locs, dirs = synth.random_los(70001, seed=5)
    los = G.ctx.los_from_MSO(locs, dirs)
    os.environ["B200RT_SCRATCH_BYTES"] = str(8 << 20)               # ~2e4 lines of sight per batch
    os.environ["B200RT_LOS_ORDER_MIN"] = "1000"                     # ... each processed longest-first
    try:
        a = G.ctx.brightness(los, 6)
    finally:
        del os.environ["B200RT_SCRATCH_BYTES"], os.environ["B200RT_LOS_ORDER_MIN"]
    assert G.ctx.kernel_ms(binding.PH_BRIGHTNESS)[1] >= 4           # several batches: one march launch each ...
    assert G.ctx.kernel_ms(binding.PH_ORDER)[1] == 3 * G.ctx.kernel_ms(binding.PH_BRIGHTNESS)[1]   # ... + 3 ordering launches
    G.ctx.los_upload(los)
    G.ctx.brightness_resident(6)                                    # one batch, input order (70001 < the order threshold)
    assert G.ctx.kernel_ms(binding.PH_BRIGHTNESS)[1] == 1
    b = G.ctx.los_download()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    # the lines of sight stay resident after the host-buffer call
    G.ctx.brightness_resident(6)
    c = G.ctx.los_download()
    assert np.array_equal(a["brightness"], c["brightness"])


@pytest.mark.parametrize("n_los", [3000, 12000])
def test_emission_split_brightness_is_the_same_arithmetic(synth, binding, n_los):
    """two emissions and few lines of sight: launch_brightness gives every (line of sight, emission) pair its own 4-lane
    group (brightness.cu, SPLIT) instead of one group per line of sight carrying both emissions.  Bit-identical results,
    with the launch in input order (3000: fits the machine at once) and longest-first (12000: 24000 items > 18944 groups)"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2, sza_T_contrast=0.1)
    G = binding.GpuModel(scn, "f64")
    for e in range(2):
        G.set_sourcefn(e, np.linspace(1.0, 0.05, scn.n_vox) * (1 + e))
    locs, dirs = synth.random_los(n_los, seed=11)
    G.ctx.los_upload(G.ctx.los_from_MSO(locs, dirs))
    G.ctx.brightness_resident(10)
    n_order = G.ctx.kernel_ms(binding.PH_ORDER)[1]
    assert n_order == (3 if n_los == 12000 else 0)
    steps = G.ctx.last_substep_count()
    assert steps > 0
    a = G.ctx.los_download()
    os.environ["B200RT_EM_SPLIT_MAX"] = "0"
    try:
        G.ctx.brightness_resident(10)
    finally:
        del os.environ["B200RT_EM_SPLIT_MAX"]
    assert G.ctx.kernel_ms(binding.PH_ORDER)[1] == 0
    assert G.ctx.last_substep_count() == steps        # sub-steps are counted once per line of sight either way
    b = G.ctx.los_download()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.any(a["brightness"][1] != a["brightness"][0])


def test_edge_cases(synth, binding, oraclebind):
    scn = synth.make_scenario(12, 8, 5, 6, n_em=2)
    G = binding.GpuModel(scn, "f64")
    O = oraclebind.OracleModel(scn, "f64")
    O.build_rows(); O.solve()
    for e in range(2):
        G.set_sourcefn(e, O.vectors(e)["S"])
    rM = synth.rMars
    locs = np.array([[30 * rM, 0, 0],          # looks away from the planet: misses the grid
                     [30 * rM, 0, 0],          # straight at the planet: exits through the bottom
                     [0, 0, 2 * rM],           # inside the grid looking up
                     [0, 0, 2 * rM]])          # inside the grid looking down
    dirs = np.array([[1.0, 0, 0], [-1.0, 0, 0], [0, 0, 1.0], [0, 0, -1.0]])
    _, bg = G.brightness(locs, dirs, 10)
    _, bo = O.brightness(locs, dirs, 10)
    assert (bg[:, :, 0] == 0).all() and (bg[:, 2, 0] == 0).all()      # miss: brightness 0, tau 0
    assert (bg[:, 2, 1] == -1).all() and (bg[:, 2, 3] == -1).all()    # planet hits: tau_absorber_final = -1
    assert (bg[:, 2, 2] >= 0).all()
    for q in range(4):
        assert rel_err(bo[:, q], bg[:, q], floor=1e-300) < 1e-6
    # error behaviour: n_subsamples = 1 is illegal (RT_grid.hpp:237), empty observation (RT_grid.hpp:302)
    los = G.ctx.los_from_MSO(locs, dirs)
    with pytest.raises(binding.B200RTError):
        G.ctx.brightness(los, 1)
    with pytest.raises(binding.B200RTError):
        G.ctx.brightness([a[:0] for a in los], 10)
    # brightness before any source function is a state error, not garbage
    G3 = binding.GpuModel(scn, "f64")
    with pytest.raises(binding.B200RTError):
        G3.ctx.brightness(los, 10)
    with pytest.raises(binding.B200RTError):
        G3.ctx.solve()


def test_layout_flag(synth, binding):
    scn = synth.make_scenario(8, 6, 4, 4, n_em=1)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    Kr = G.ctx.influence_matrix(0, binding.ROW_MAJOR)
    Kc = G.ctx.influence_matrix(0, binding.COL_MAJOR)
    assert np.array_equal(Kr, Kc.T)


def test_not_dominant_is_reported(synth, binding):
    """a branching ratio that breaks row dominance must come back as a status, not a wrong answer"""
    scn = synth.make_scenario(8, 6, 4, 4, n_em=1)
    scn.em_scalars[0][0] = 50.0
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    with pytest.raises(binding.B200RTError, match="dominant"):
        G.solve()


def test_traversal_special_lines_of_sight(synth, binding, oraclebind):
    """hand-made lines of sight through the corners of the fast traversal path: origins on the Mars-Sun axis, exactly
    on a radial boundary and on a voxel point; directions exactly radial (out / in), exactly along +-z and tangent to a
    boundary sphere; origins below the grid and beyond it.  Lists must be the oracle's, bit for bit, in both
    precisions (rays the fast path cannot verify fall back to the exact ranking)."""
    for shape in ((12, 8, 5, 6), (40, 20, 7, 12)):
        scn = synth.make_scenario(*shape, n_em=1)
        rb = np.asarray(scn.rb)
        pts_r = np.sqrt(rb[:-1] * rb[1:])
        locs, dirs = [], []

        def add(p, d):
            d = np.asarray(d, dtype=np.float64)
            locs.append(np.asarray(p, dtype=np.float64))
            dirs.append(d / np.linalg.norm(d))
        for r in (pts_r[1], rb[2], pts_r[len(pts_r) // 2], rb[-2], pts_r[-1], 0.9 * rb[0], 1.5 * rb[-1]):
            for p in ([r, 0, 0], [-r, 0, 0], [0, r, 0], [0, 0, r], [r / np.sqrt(2), 0, r / np.sqrt(2)], [0.3 * r, -0.5 * r, np.sqrt(1 - 0.34) * r]):
                p = np.asarray(p, dtype=np.float64)
                n = p / np.linalg.norm(p)
                t = np.cross(n, [0.3, 0.7, 0.2]); t /= np.linalg.norm(t)
                for d in (n, -n, t, -t, [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, -1, 0], n + 1e-9 * t, t + 1e-12 * n, 0.6 * n + 0.8 * t):
                    add(p, d)
        locs, dirs = np.array(locs), np.array(dirs)
        for prec in ("f64", "f32"):
            O = oraclebind.OracleModel(scn, prec)
            G = binding.GpuModel(scn, prec)
            a, b = O.traverse_los(locs, dirs), G.traverse_los(locs, dirs)
            assert_lists_equal(a[:4], b[:4])
            assert (a[0] == 0).any() and (a[0] > 8).any()           # misses and long lists both occur
