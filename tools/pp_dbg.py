import sys, importlib, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
synth = importlib.import_module('3d_planetary_rt_model_b200.synth')
binding = importlib.import_module('3d_planetary_rt_model_b200.binding')
from oracle import oraclebind
np.set_printoptions(linewidth=200, precision=4)
for prec in ('f64', 'f32'):
    for shape in ((8, 4), (12, 5), (40, 7)):
        scn = synth.make_scenario_pp(*shape, n_em=2)
        O = oraclebind.OracleModel(scn, prec); G = binding.GpuModel(scn, prec)
        a, b = O.traverse_voxel_rays(), G.traverse_voxel_rays()
        print(prec, shape, 'lists equal', all(np.array_equal(x, y) for x, y in zip(a, b)), 'max dist', a[3].max())
        O.build_rows(); G.build_rows()
        for e in range(2):
            Ko, Kg = O.K(e), G.K(e)
            bad = ~np.isfinite(Kg)
            err = np.abs(Ko - Kg) / np.maximum(np.abs(Ko), 1e-300)
            err[bad] = np.inf
            i, j = np.unravel_index(np.argmax(err), err.shape)
            print('  em', e, 'nonfinite', bad.sum(), 'worst', (i, j), Ko[i, j], Kg[i, j], 'relerr', err[i, j])
            vo, vg = O.vectors(e), G.vectors(e, want_S=False)
            for k in ("S0", "tau_species_ss", "tau_absorber_ss"):
                print('   ', k, np.abs(vo[k] - vg[k]).max() / np.abs(vo[k]).max())
