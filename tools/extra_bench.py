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


def iph(n_los):
    synth = importlib.import_module(PKG + ".synth")
    binding = importlib.import_module(PKG + ".binding")
    ctx = binding.Context(0, binding.F64)
    ctx.iph_set_table(synth.make_iph_table())
    ra, dec = synth.random_sky(n_los)
    g, pos = synth.lyman_alpha_typical_g_factor, synth.MARS_ECLIPTIC_POS
    ctx.iph_model(g, pos, ra[:1000], dec[:1000])                       # warm-up
    t0 = time.perf_counter()
    out = ctx.iph_model(g, pos, ra, dec)
    wall = time.perf_counter() - t0
    ms, _ = ctx.kernel_ms(binding.PH_IPH)
    return {"iph_n_los": n_los, "iph_kernel_ms": ms, "iph_los_per_s_kernel": n_los / (ms * 1e-3),
            "iph_los_per_s_e2e": n_los / wall, "iph_mean_kR": float(out.mean())}


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
    t0 = time.perf_counter()
    b = F.brightness_batch(NH, TT, contexts, gpus)
    wall = time.perf_counter() - t0
    return {"sweep_sets": len(NH), "sweep_los_per_set": n_los, "sweep_contexts_per_gpu": contexts,
            "sweep_seconds": wall, "sweep_sets_per_s": len(NH) / wall, "sweep_finite": bool(np.isfinite(b).all())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iph-los", type=int, default=1000000)
    ap.add_argument("--sets", type=int, default=512)
    ap.add_argument("--sweep-los", type=int, default=10000)
    ap.add_argument("--contexts", type=int, default=4)
    ap.add_argument("--gpus", type=int, default=-1)
    a = ap.parse_args()
    out = {}
    if a.iph_los > 0:
        out.update(iph(a.iph_los))
    if a.sets > 0:
        out.update(sweep(a.sets, a.sweep_los, a.contexts, a.gpus))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
