"""IPH background and (nH, T) sweep throughput on the GPU(s) of this box (BASELINE.json configs[2] IPH part, configs[3]).
python tools/extra_bench.py [--iph-los 1000000] [--sets 512] [--sweep-los 10000] [--contexts 4] [--gpus -1]"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "3d_planetary_rt_model_b200"


def real_iph_table():
    """the reference's Quemerais table (tests/golden/iph_real_table.npz holds its values); the synthetic table of
    synth.make_iph_table where the fixture is absent"""
    p = os.path.join(ROOT, "tests", "golden", "iph_real_table.npz")
    if not os.path.exists(p):
        return importlib.import_module(PKG + ".synth").make_iph_table(), "synthetic"
    z = np.load(p)
    tab = {k[4:]: z[k] for k in z.files if k.startswith("tab_")}
    for k in ("kmax", "lmax", "ninf"):
        tab[k] = int(tab[k])
    tab["temp"] = float(tab["temp"])
    return tab, "reference table (fsm99td12v20t80)"


def iph(n_los, n_gpus=1):
    """IPH background of n_los lines of sight, split over n_gpus devices of this process (one handle)"""
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    ctx = binding.Context(0, binding.F64) if n_gpus <= 1 else binding.Context(precision=binding.F64, devices=list(range(n_gpus)))
    tab, which = real_iph_table()
    ctx.iph_set_table(tab)
    ra, dec = synth.random_sky(n_los)
    g, pos = synth.lyman_alpha_typical_g_factor, synth.MARS_ECLIPTIC_POS
    ctx.iph_model(g, pos, ra[:1000 * max(1, n_gpus)], dec[:1000 * max(1, n_gpus)])   # warm-up
    ctx.iph_model(g, pos, ra, dec)
    t0 = time.perf_counter()
    out = ctx.iph_model(g, pos, ra, dec)
    wall = time.perf_counter() - t0
    ms, _ = ctx.kernel_ms(binding.PH_IPH)
    return {"iph_n_los": n_los, "iph_n_gpus": n_gpus, "iph_table": which, "iph_kernel_ms": ms,
            "iph_los_per_s_kernel": n_los / (ms * 1e-3), "iph_los_per_s_e2e": n_los / wall, "iph_mean_kR": float(out.mean())}


def sweep(n_sets, n_los, contexts, gpus):
    synth = importlib.import_module(PKG + ".synth")
    hb = importlib.import_module(PKG + ".host_binding")
    F = hb.Pyobservation_fit()
    locs, dirs = synth.random_los(n_los)
    F.add_observation(locs, dirs)
    nn = int(round(n_sets ** 0.5 * (2 ** 0.5)))                        # 32 x 16 for 512
    nH = np.logspace(4, 7, nn)
    T = np.linspace(100, 400, max(1, n_sets // nn))
    NH, TT = np.meshgrid(nH, T, indexing="ij")
    NH, TT = NH.ravel()[:n_sets], TT.ravel()[:n_sets]
    F.brightness_batch(NH[:8], TT[:8], contexts, gpus)                 # warm-up (context creation, first launches)
    # a set is a few hundred microseconds of kernels between a dozen host round trips, so the rate follows the host's
    # scheduling noise (other tenants of the box): three runs, the best one is reported next to all of them
    walls = []
    for _ in range(3):
        t0 = time.perf_counter()
        b = F.brightness_batch(NH, TT, contexts, gpus)
        walls.append(time.perf_counter() - t0)
    wall = min(walls)
    return {"sweep_sets": len(NH), "sweep_los_per_set": n_los, "sweep_contexts_per_gpu": contexts, "sweep_n_gpus": gpus,
            "sweep_seconds": wall, "sweep_sets_per_s": len(NH) / wall, "sweep_sets_per_s_runs": [len(NH) / w for w in walls],
            "sweep_finite": bool(np.isfinite(b).all())}


def multiplet(kind, image_px=600):
    """BASELINE.json configs[4] (ii)/(iii): one multiplet emission on the reference default grid 40x20x7x12 --
    source function (influence + single scattering + solve) and the 600 x 600 fake image (observation::fake,
    observation.hpp:173-208; generate_source_function.cpp:296-306), device times of the kernels"""
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    scn = synth.make_multiplet_scenario(kind, 40, 20, 7, 12)
    G = binding.GpuMultiplet(scn, "f64")
    locs, dirs = synth.fake_image(30 * synth.rMars, 30, image_px)
    los = G.ctx.los_from_MSO(locs, dirs)
    G.ctx.los_upload(los)
    out = {}
    for it in range(3):
        G.ctx.influence()
        t_tr, t_in = G.ctx.kernel_ms(binding.PH_TRAVERSE)[0], G.ctx.kernel_ms(binding.PH_INFLUENCE)[0]
        steps = G.ctx.last_step_count()
        G.ctx.solve()
        t_so = G.ctx.kernel_ms(binding.PH_SOLVE)[0]
        G.ctx.brightness_resident(10)
        t_lt, t_br = G.ctx.kernel_ms(binding.PH_TRAVERSE)[0], G.ctx.kernel_ms(binding.PH_BRIGHTNESS)[0]
    name = {0: "O1026", 1: "H_lyman_multiplet", 2: "H_lyman_singlet"}[kind]
    b = G.ctx.los_download()["brightness"]
    out[f"mult_{name}"] = {"grid": "40x20x7x12", "n_elements": G.n_el, "ray_voxel_steps": steps,
                           "influence_traverse_ms": t_tr, "influence_march_ms": t_in,
                           "steps_per_s": steps / ((t_tr + t_in) * 1e-3), "solve_ms": t_so, "residual": G.ctx.residual(0),
                           "n_los": len(locs), "los_traverse_ms": t_lt, "brightness_ms": t_br,
                           "los_per_s": len(locs) / ((t_lt + t_br) * 1e-3), "max_line_kR": [float(x) for x in b.max(axis=1)]}
    return out


def float_job(n_los):
    """The bench workload with Real = float (the only precision the reference's own GPU module is built in, makefile:80,230;
    parity bar 1e-4): 100x60 grid, 24x16 rays, H Ly alpha + n_los lines of sight; the solve stays FP64."""
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    scn = synth.make_scenario(100, 60, 24, 16, n_em=1, rmethod=synth.RMETHOD_ALTITUDE, rmax=synth.rMars + 50000e5)
    G = binding.GpuModel(scn, "f32")
    locs, dirs = synth.random_los(n_los)
    G.ctx.los_upload(G.ctx.los_from_MSO(locs, dirs))
    for it in range(3):
        G.ctx.influence(0, scn.n_vox)
        t_tr, t_in = G.ctx.kernel_ms(binding.PH_TRAVERSE)[0], G.ctx.kernel_ms(binding.PH_INFLUENCE)[0]
        steps = G.ctx.last_step_count()
        G.ctx.solve()
        t_so = G.ctx.kernel_ms(binding.PH_SOLVE)[0]
        G.ctx.brightness_resident(10)
        t_lt, t_br = G.ctx.kernel_ms(binding.PH_TRAVERSE)[0], G.ctx.kernel_ms(binding.PH_BRIGHTNESS)[0]
    return {"f32_job": {"influence_traverse_ms": t_tr, "influence_march_ms": t_in, "steps_per_s": steps / ((t_tr + t_in) * 1e-3),
                        "solve_ms": t_so, "los_traverse_ms": t_lt, "brightness_ms": t_br,
                        "los_per_s": n_los / ((t_lt + t_br) * 1e-3), "job_ms": t_tr + t_in + t_so + t_lt + t_br}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iph-los", type=int, default=1000000)
    ap.add_argument("--sets", type=int, default=512)
    ap.add_argument("--sweep-los", type=int, default=10000)
    ap.add_argument("--contexts", type=int, default=4)
    ap.add_argument("--gpus", type=int, default=-1)
    ap.add_argument("--multiplet", action="store_true", help="O I 102.6 and H Lyman multiplet on the default grid")
    a = ap.parse_args()
    out = {}
    if a.iph_los > 0:
        out.update(iph(a.iph_los, max(1, a.gpus)))
    if a.sets > 0:
        out.update(sweep(a.sets, a.sweep_los, a.contexts, a.gpus))
    if a.multiplet:
        out.update(multiplet(0))
        out.update(multiplet(1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
