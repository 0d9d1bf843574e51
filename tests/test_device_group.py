"""One handle, several devices of one process (b200rt_create_multi, include/b200rt.h; csrc/device_group.cu).

The N-device result must equal the one-device result: same traversal, same step counts, K / S0 / S / brightness to
rounding (the order of the fp64 REDs into K differs from run to run even on one device).  The group is built on every
visible GPU; on a one-GPU box the same device is named several times, which runs the whole fan-out -- interleaved row
shards, row sinks, source-function hand-over, line-of-sight slices landing at their offsets -- through the same code.
Under the driver's 8-GPU step the members are eight different B200s."""
import os

import numpy as np
import pytest

from util import assert_lists_equal, rel_err

pytestmark = pytest.mark.gpu


def group_devices(binding, at_least=3):
    n = binding.load().b200rt_device_count()
    ids = list(range(n))
    while len(ids) < at_least:
        ids.append(0)
    return ids


@pytest.fixture()
def split_everything(monkeypatch):
    """tiny test problems still fan out"""
    monkeypatch.setenv("B200RT_GROUP_MIN_RAYS", "0")
    monkeypatch.setenv("B200RT_GROUP_MIN_LOS", "0")


def test_plain_handle_for_one_device(binding):
    c = binding.Context(precision=binding.F64, devices=[0])
    assert c.n_devices == 1
    c.close()


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_group_equals_single_device(synth, binding, split_everything, prec):
    scn = synth.make_scenario(20, 12, 6, 8, n_em=2, sza_T_contrast=0.1)
    one = binding.GpuModel(scn, prec, device=0)
    ids = group_devices(binding)
    grp = binding.GpuModel(scn, prec, devices=ids)
    assert grp.ctx.n_devices == len(ids)
    tol = 1e-12 if prec == "f64" else 1e-5
    assert_lists_equal(one.traverse_voxel_rays(), grp.traverse_voxel_rays())
    _, s1 = one.build_rows()
    _, sN = grp.build_rows()
    assert s1 == sN
    for e in range(2):
        assert rel_err(one.K(e), grp.K(e)) < tol
        a, b = one.vectors(e, want_S=False), grp.vectors(e, want_S=False)
        for k in ("S0", "tau_species_ss", "tau_absorber_ss"):
            assert np.array_equal(a[k], b[k]), k
    r1, rN = one.solve(), grp.solve()
    assert max(rN) < 1e-12
    for e in range(2):
        assert rel_err(one.vectors(e)["S"], grp.vectors(e)["S"]) < 1e-10
        grp.set_sourcefn(e, one.vectors(e)["S"])      # identical S on every member from here on
    locs, dirs = synth.random_los(1003, seed=5)       # odd count: ragged slices
    assert_lists_equal(one.traverse_los(locs, dirs)[:4], grp.traverse_los(locs, dirs)[:4])
    for nsub in (10, 0):
        _, b1 = one.brightness(locs, dirs, nsub)      # host-buffer call: slices land at their offsets
        _, bN = grp.brightness(locs, dirs, nsub)
        assert np.array_equal(b1, bN), nsub
    # the three-step form (resident lines of sight)
    los = one.ctx.los_from_MSO(locs, dirs)
    one.ctx.los_upload(los); one.ctx.brightness_resident(10)
    grp.ctx.los_upload(los); grp.ctx.brightness_resident(10)
    d1, dN = one.ctx.los_download(), grp.ctx.los_download()
    for k in d1:
        assert np.array_equal(d1[k], dN[k]), k
    assert one.ctx.last_substep_count() == grp.ctx.last_substep_count()


def test_group_generate_S_shares_the_source_function(synth, binding, split_everything):
    """generate_S on the handle: rows from every member, solve on the first, S on all (brightness right after)"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    one = binding.GpuModel(scn, "f64", device=0)
    grp = binding.GpuModel(scn, "f64", devices=group_devices(binding, at_least=4))
    one.ctx.generate_S()
    grp.ctx.generate_S()
    assert grp.ctx.kernel_ms(binding.PH_SOLVE)[0] > 0 and grp.ctx.kernel_ms(binding.PH_INFLUENCE)[0] > 0
    locs, dirs = synth.random_los(500, seed=3)
    _, b1 = one.brightness(locs, dirs, 10)
    _, bN = grp.brightness(locs, dirs, 10)
    assert rel_err(b1[:, 0], bN[:, 0]) < 1e-9


def test_group_keeps_lines_of_sight_across_a_regrid(synth, binding, split_everything):
    """resident lines of sight survive b200rt_set_grid_* on a plain context (observation_fit re-grids on every parameter
    set and uploads its observations once); the group handle must behave the same"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    locs, dirs = synth.random_los(700, seed=8)
    res = []
    for devices in (None, group_devices(binding)):
        M = binding.GpuModel(scn, "f64", device=0, devices=devices)
        M.ctx.los_upload(M.ctx.los_from_MSO(locs, dirs))
        M.ctx.generate_S()
        M.ctx.brightness_resident(10)
        first = M.ctx.los_download()["brightness"].copy()
        M.ctx.set_grid(M.g)                                    # re-grid + new tables, no new upload
        b, T, s, g = (float(x) for x in scn.em_scalars[0])
        M.ctx.set_singlet(0, 1, b, T, s, g, binding.define_singlet_tables(scn, 0))
        M.ctx.generate_S()
        M.ctx.brightness_resident(10)
        assert rel_err(first, M.ctx.los_download()["brightness"]) < 1e-9
        res.append(first)
    assert rel_err(res[0], res[1]) < 1e-9          # (S differs in the last bits: the order of the REDs into K)


def test_group_small_work_stays_on_first_device(synth, binding):
    """default thresholds: a 12x8 grid and 500 lines of sight are not worth the fan-out"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    grp = binding.GpuModel(scn, "f64", devices=group_devices(binding))
    one = binding.GpuModel(scn, "f64", device=0)
    grp.ctx.generate_S(); one.ctx.generate_S()
    assert grp.ctx.kernel_ms(binding.PH_TRAVERSE)[1] == one.ctx.kernel_ms(binding.PH_TRAVERSE)[1]   # one member's launches
    locs, dirs = synth.random_los(500, seed=3)
    assert rel_err(one.brightness(locs, dirs, 10)[1], grp.brightness(locs, dirs, 10)[1]) < 1e-9


def test_group_multiplet(synth, binding, split_everything):
    """multiplet emissions: rows on the first device (no row sink for them), lines of sight split"""
    scn = synth.make_multiplet_scenario(0, 10, 6, 4, 4)
    one = binding.GpuMultiplet(scn, "f64", device=0)
    grp = binding.GpuMultiplet(scn, "f64", devices=group_devices(binding))
    for m in (one, grp):
        m.build_rows()
        assert m.solve() < 1e-12
    assert rel_err(one.vectors()["S"], grp.vectors()["S"]) < 1e-10
    grp.set_sourcefn(one.vectors()["S"])
    locs, dirs = synth.random_los(301, seed=9)
    b1, bN = one.brightness(locs, dirs, 10), grp.brightness(locs, dirs, 10)
    for k in b1:
        assert np.array_equal(b1[k], bN[k]), k


def test_group_iph(synth, binding, split_everything):
    one = binding.Context(0, binding.F64)
    grp = binding.Context(precision=binding.F64, devices=group_devices(binding))
    tab = synth.make_iph_table()
    one.iph_set_table(tab); grp.iph_set_table(tab)
    ra, dec = synth.random_sky(777)
    g, pos = synth.lyman_alpha_typical_g_factor, synth.MARS_ECLIPTIC_POS
    assert np.array_equal(one.iph_model(g, pos, ra, dec), grp.iph_model(g, pos, ra, dec))


def test_facade_uses_every_device(synth, binding, split_everything, monkeypatch):
    """observation_fit::generate_source_function + brightness on the handle (the reference's user calls,
    observation_fit.cpp:122-169,491-516) == the same calls on one device"""
    import importlib
    hb = importlib.import_module("3d_planetary_rt_model_b200.host_binding")
    locs, dirs = synth.random_los(400, seed=2)
    out = []
    for dev in (0, -1):
        if dev < 0:
            monkeypatch.setenv("B200RT_DEVICES", ",".join(str(i) for i in group_devices(binding)))
        F = hb.Pyobservation_fit(device=dev)
        F.add_observation(locs, dirs)
        F.generate_source_function(1e5, 300.0)
        out.append(np.asarray(F.brightness()))
    assert rel_err(out[0], out[1]) < 1e-9


@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_group_distributed_solve(synth, binding, split_everything, monkeypatch, prec):
    """grids of B200RT_KRYLOV_MIN_N voxels and more: the rows stay on the members that built them and the members solve
    together (csrc/solve_krylov.cu); S is resident on every member, and a caller that asks for the assembled influence
    matrix afterwards still gets every row (gathered on demand)"""
    monkeypatch.setenv("B200RT_KRYLOV_MIN_N", "50")
    scn = synth.make_scenario(20, 12, 6, 8, n_em=2, sza_T_contrast=0.1)
    one = binding.GpuModel(scn, prec, device=0)
    one.build_rows()
    one.solve()
    ids = group_devices(binding)
    grp = binding.GpuModel(scn, prec, devices=ids)
    grp.ctx.generate_S()
    steps = grp.ctx.last_solve_steps()
    assert 5 < steps < 120                                        # it was the distributed solve
    tol = 1e-7 if prec == "f64" else 1e-5                          # float tables: K itself differs by float rounding order
    for e in range(2):
        assert grp.ctx.residual(e) < 1e-12
        assert rel_err(one.vectors(e)["S"], grp.vectors(e)["S"], floor=1e-30) < tol
    locs, dirs = synth.random_los(1003, seed=5)
    _, bN = grp.brightness(locs, dirs, 10)                         # every member integrates with its resident S
    for e in range(2):
        one.set_sourcefn(e, grp.vectors(e)["S"])
    _, b1 = one.brightness(locs, dirs, 10)
    assert rel_err(b1, bN, floor=1e-300) < (1e-12 if prec == "f64" else 1e-5)
    for e in range(2):
        assert rel_err(one.K(e), grp.K(e)) < (1e-12 if prec == "f64" else 1e-5)     # gathered from the members' shards
    grp.ctx.solve()                                                # again, on the same rows
    assert grp.ctx.last_solve_steps() == steps
