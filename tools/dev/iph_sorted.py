"""Development aid: IPH kernel time for lines of sight in random order vs grouped by direction"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
synth = importlib.import_module("3d_planetary_rt_model_b200.synth")
binding = importlib.import_module("3d_planetary_rt_model_b200.binding")
ctx = binding.Context(0, binding.F64)
ctx.iph_set_table(synth.make_iph_table())
n = 1000000
ra, dec = synth.random_sky(n)
g, pos = synth.lyman_alpha_typical_g_factor, synth.MARS_ECLIPTIC_POS
ctx.iph_model(g, pos, ra[:1000], dec[:1000])
out = ctx.iph_model(g, pos, ra, dec); print("random order   ", ctx.kernel_ms(binding.PH_IPH)[0], "ms")
# group by direction: HEALPix-like coarse cells then fine order
x = np.cos(np.radians(dec)) * np.cos(np.radians(ra)); y = np.cos(np.radians(dec)) * np.sin(np.radians(ra)); z = np.sin(np.radians(dec))
for nb in (16, 64, 256):
    key = (np.floor((z + 1) / 2 * nb).astype(np.int64) * 4096 + np.floor((np.arctan2(y, x) + np.pi) / (2 * np.pi) * 4095).astype(np.int64))
    o = np.argsort(key, kind="stable")
    out2 = ctx.iph_model(g, pos, ra[o], dec[o]); print(f"sorted (nb={nb}) ", ctx.kernel_ms(binding.PH_IPH)[0], "ms", "same values:", np.array_equal(out2, out[o]))
