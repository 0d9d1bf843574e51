"""b200rt_solve_distributed (csrc/solve_krylov.cu): GMRES on (I - w K) S = S0 with the rows of K left on the ranks
that built them, against the dense LU of the same library and against the oracle's solve.

One GPU is enough to exercise the exchange protocol: several contexts on device 0, one host thread each, every
context holding only ITS interleaved shard of the rows and writing its pieces into all the exchange blocks (plain
device pointers inside one process).  The 2-GPU form of the same test runs when the box has two devices.
"""
import importlib
import os
import threading

import numpy as np
import pytest

from util import rel_err

pytestmark = pytest.mark.gpu

multi = importlib.import_module("3d_planetary_rt_model_b200.multi")


def lu_solution(binding, scn):
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    res = G.solve()
    return G, [G.vectors(e)["S"].copy() for e in range(scn.n_em)], res


def run_ranks(binding, scn, devices, chunks=3, prec="f64"):
    """one context per entry of `devices`, each building its interleaved shard and joining the distributed solve"""
    world = len(devices)
    models = [binding.GpuModel(scn, prec, device=d) for d in devices]
    blocks = [m.ctx.solve_exchange()[0] for m in models]
    errors = [None] * world

    def work(r):
        try:
            ranges = multi.partition_interleaved(scn.n_vox, world, r, chunks)
            models[r].ctx.influence_ranges(ranges)
            models[r].ctx.solve_distributed(r, world, blocks)
        except Exception as ex:          # noqa: BLE001 -- reported by the caller
            errors[r] = ex

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for ex in errors:
        if ex is not None:
            raise ex
    return models


@pytest.fixture(params=["one launch", "four launches per step"])
def form(request, monkeypatch):
    """the solve as one cooperative launch (kry_loop, the default) and as the per-step launches it falls back to"""
    monkeypatch.setenv("B200RT_KRYLOV_FUSED", "1" if request.param == "one launch" else "0")
    monkeypatch.setenv("B200RT_KRYLOV_CTAS", "64")     # several ranks share device 0 here: every resident grid has to fit
    return request.param


@pytest.mark.parametrize("shape", [(12, 8, 5, 6), (40, 20, 7, 12)])
def test_one_rank_gmres_equals_lu_and_oracle(synth, binding, oraclebind, shape, form):
    scn = synth.make_scenario(*shape, n_em=2, sza_T_contrast=0.1)
    G, S_lu, _ = lu_solution(binding, scn)
    block, _ = G.ctx.solve_exchange()
    G.ctx.solve_distributed(0, 1, [block])
    steps = G.ctx.last_solve_steps()
    assert 5 < steps < 120
    O = oraclebind.OracleModel(scn, "f64")
    O.build_rows()
    O.solve()
    sol = [G.vectors(e)["S"].copy() for e in range(2)]
    for e in range(2):
        S = sol[e]
        assert G.ctx.residual(e) < 1e-12                   # the true residual, from one more product with K
        assert rel_err(S_lu[e], S, floor=1e-30) < 1e-7      # element by element, against the 1e-6 bar
        assert rel_err(O.vectors(e)["S"], S, floor=1e-30) < 1e-6
    # a second call continues the round counters of the same block
    G.ctx.solve_distributed(0, 1, [block])
    assert G.ctx.last_solve_steps() == steps
    for e in range(2):
        assert np.array_equal(G.vectors(e)["S"], sol[e])


@pytest.mark.parametrize("world", [2, 3])
def test_ranks_sharing_one_device(synth, binding, world, form):
    """the exchange protocol itself: `world` contexts on device 0, each with its own rows only"""
    scn = synth.make_scenario(20, 12, 6, 8, n_em=2, sza_T_contrast=0.1)
    _, S_lu, _ = lu_solution(binding, scn)
    models = run_ranks(binding, scn, [0] * world)
    ref = [models[0].vectors(e)["S"] for e in range(2)]
    for e in range(2):
        assert rel_err(S_lu[e], ref[e], floor=1e-30) < 1e-7
    for m in models[1:]:
        assert m.ctx.last_solve_steps() == models[0].ctx.last_solve_steps()
        for e in range(2):
            assert np.array_equal(m.vectors(e)["S"], ref[e])      # every rank ran the same arithmetic
            assert m.ctx.residual(e) == models[0].ctx.residual(e)
    # and the brightness of a rank's lines of sight uses that resident S
    locs, dirs = synth.random_los(500, seed=2)
    los = models[1].ctx.los_from_MSO(locs, dirs)
    a = models[1].ctx.brightness(los, 6)
    assert np.isfinite(a["brightness"]).all()


def test_missing_rows_are_reported(synth, binding, form):
    """a rank that built fewer rows than its share: the iteration cannot converge, and says so instead of hanging"""
    scn = synth.make_scenario(12, 8, 5, 6, n_em=1)
    G = binding.GpuModel(scn, "f64")
    G.ctx.influence_ranges([(0, scn.n_vox // 2)])
    block, _ = G.ctx.solve_exchange()
    with pytest.raises(binding.B200RTError):
        G.ctx.solve_distributed(0, 1, [block])


def test_both_forms_agree(synth, binding, monkeypatch):
    """one cooperative launch against four launches per step: same algorithm and order of sums (they differ in where
    1 / |w| multiplies the product, i.e. by rounding)"""
    scn = synth.make_scenario(20, 12, 6, 8, n_em=1)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    block, _ = G.ctx.solve_exchange()
    out = []
    for fused in ("1", "0"):
        monkeypatch.setenv("B200RT_KRYLOV_FUSED", fused)
        G.ctx.solve_distributed(0, 1, [block])
        out.append((G.vectors(0)["S"].copy(), G.ctx.last_solve_steps(), G.ctx.residual(0), G.ctx.kernel_ms(binding.PH_SOLVE)[1]))
    assert out[0][3] <= 5 and out[1][3] > 20                        # launches: one (+ four for the preconditioner), against four per step
    assert out[0][1] <= out[1][1]                                   # the one-launch form iterates on the preconditioned system
    assert rel_err(out[0][0], out[1][0], floor=1e-30) < 1e-7
    assert out[0][2] < 1e-12 and out[1][2] < 1e-12


def test_two_devices(synth, binding):
    if binding.load().b200rt_device_count() < 2:
        pytest.skip("one visible device")
    scn = synth.make_scenario(40, 20, 7, 12, n_em=1)
    _, S_lu, _ = lu_solution(binding, scn)
    G = binding.GpuModel(scn, "f64", devices=[0, 1])                 # the group enables peer access between its devices
    del G
    models = run_ranks(binding, scn, [0, 1], chunks=8)
    for m in models:
        assert rel_err(S_lu[0], m.vectors(0)["S"], floor=1e-30) < 1e-7
    assert np.array_equal(models[0].vectors(0)["S"], models[1].vectors(0)["S"])


def test_b200rt_solve_picks_by_size_and_falls_back(synth, binding, monkeypatch):
    """b200rt_solve on a context that built every row: GMRES from B200RT_KRYLOV_MIN_N unknowns, the LU below -- and the LU
    again when the iteration is cut short before it converges"""
    scn = synth.make_scenario(20, 12, 6, 8, n_em=1)
    G = binding.GpuModel(scn, "f64")
    G.build_rows()
    G.solve()                                                        # 209 unknowns: the LU
    S_lu = G.vectors(0)["S"].copy()
    assert G.ctx.kernel_ms(binding.PH_SOLVE)[1] > 5                  # its launches
    monkeypatch.setenv("B200RT_KRYLOV_MIN_N", "50")
    G.solve()                                                        # now the iteration (one launch + the set-up)
    assert G.ctx.kernel_ms(binding.PH_SOLVE)[1] <= 5 and 3 < G.ctx.last_solve_steps() < 100
    assert rel_err(S_lu, G.vectors(0)["S"], floor=1e-30) < 1e-9 and G.ctx.residual(0) < 1e-12
    monkeypatch.setenv("B200RT_KRYLOV_MAXIT", "3")
    G.solve()                                                        # cut short -> not converged -> the LU decides
    assert G.ctx.last_solve_steps() == 3
    assert np.array_equal(G.vectors(0)["S"], S_lu) and G.ctx.residual(0) < 1e-12
    monkeypatch.setenv("B200RT_SOLVER", "lu")
    monkeypatch.delenv("B200RT_KRYLOV_MAXIT")
    G.solve()
    assert np.array_equal(G.vectors(0)["S"], S_lu)
