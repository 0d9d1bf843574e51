"""ctypes binding of the C ABI in include/b200rt.h (libb200rt.so).

This is the same thin stub a maintainer of the reference would add on the Python
side (INTEGRATION.md shows the C++ one).  There is no fallback of any kind: if the
shared library is missing, or no CUDA device is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200rt.so")

F64, F32 = 0, 1
ROW_MAJOR, COL_MAJOR = 0, 1
PH_TRAVERSE, PH_INFLUENCE, PH_SOLVE, PH_BRIGHTNESS, PH_IPH = 0, 1, 2, 3, 4

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_vp = C.c_void_p

# every symbol include/b200rt.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "b200rt_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_vp)]),
    "b200rt_destroy": (C.c_int, [_vp]),
    "b200rt_last_error": (C.c_char_p, [_vp]),
    "b200rt_device_count": (C.c_int, []),
    "b200rt_set_grid_sph": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int] + [_dp] * 7),
    "b200rt_make_grid_sph": (C.c_int, [C.c_int] * 5 + [_dp, C.c_int, C.c_int] + [_dp] * 6),
    "b200rt_set_grid_pp": (C.c_int, [_vp, C.c_int, C.c_int] + [_dp] * 4),
    "b200rt_make_grid_pp": (C.c_int, [C.c_int] * 3 + [_dp] * 4),
    "b200rt_set_singlet": (C.c_int, [_vp, C.c_int, C.c_int] + [C.c_double] * 4 + [_dp] * 8),
    "b200rt_set_g_factor": (C.c_int, [_vp, C.c_int, C.c_double]),
    "b200rt_generate_S": (C.c_int, [_vp]),
    "b200rt_influence": (C.c_int, [_vp, C.c_int, C.c_int]),
    "b200rt_solve": (C.c_int, [_vp]),
    "b200rt_last_step_count": (C.c_int, [_vp, C.POINTER(C.c_longlong)]),
    "b200rt_get_solution": (C.c_int, [_vp] + [C.c_int] + [_vp] * 4),
    "b200rt_get_influence": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "b200rt_set_sourcefn": (C.c_int, [_vp, C.c_int, _dp]),
    "b200rt_last_residual": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    "b200rt_influence_dev": (C.c_int, [_vp, C.c_int] + [C.POINTER(_vp)] * 4),
    "b200rt_sourcefn_dev": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "b200rt_los_from_MSO": (C.c_int, [C.c_int, C.c_int] + [_dp] * 11),
    "b200rt_brightness": (C.c_int, [_vp, C.c_int] + [_dp] * 9 + [C.c_int] + [_vp] * 4),
    "b200rt_los_upload": (C.c_int, [_vp, C.c_int] + [_dp] * 9),
    "b200rt_brightness_resident": (C.c_int, [_vp, C.c_int]),
    "b200rt_los_download": (C.c_int, [_vp] + [_vp] * 4),
    "b200rt_iph_load_table": (C.c_int, [_vp, C.c_char_p]),
    "b200rt_iph_set_table": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_float] + [_fp] * 7),
    "b200rt_iph_background": (C.c_int, [_vp] + [C.c_float] * 4 + [C.c_int, _fp, _fp, _fp, _fp, _vp]),
    "b200rt_iph_model": (C.c_int, [_vp, C.c_double, _dp, C.c_int, _dp, _dp, _dp]),
    "b200rt_iph_extinction": (C.c_int, [C.c_int, _dp, _dp, _dp]),
    "b200rt_traverse_voxel_rays": (C.c_int, [_vp, C.c_int, C.c_int, C.c_longlong, _ip, _ip, _ip, _dp,
                                             C.POINTER(C.c_longlong)]),
    "b200rt_traverse_los": (C.c_int, [_vp, C.c_longlong, _ip, _ip, _ip, _dp, C.POINTER(C.c_longlong)]),
    "b200rt_last_kernel_ms": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "b200rt_synchronize": (C.c_int, [_vp]),
    "b200rt_measure_fp64_peaks": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """dlopen libb200rt.so and type every entry point; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run __graft_entry__.build() (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class B200RTError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Context:
    """One b200rt_ctx (one GPU).  Method names follow the reference's RT_grid where one exists."""

    def __init__(self, device: int = 0, precision: int = F64):
        self.lib = load()
        self.h = _vp()
        rc = self.lib.b200rt_create(device, precision, C.byref(self.h))
        if rc != 0:
            raise B200RTError(f"b200rt_create(device={device}) failed with status {rc}: no usable CUDA device "
                              "(this library has no CPU path)")
        self.precision = precision
        self.n_vox = 0
        self.n_em = 0
        self.n_los = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200rt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise B200RTError(f"status {rc}: {self.lib.b200rt_last_error(self.h).decode()}")

    # ---- geometry
    def make_grid(self, n_rb, n_sb, n_theta, n_phi, rb, szamethod, raymethod):
        n_rays = n_theta * n_phi
        out = dict(radial_boundaries=np.ascontiguousarray(rb, dtype=np.float64),
                   sza_boundaries=np.zeros(n_sb), pts_radii=np.zeros(n_rb - 1), pts_sza=np.zeros(n_sb - 1),
                   ray_theta=np.zeros(n_rays), ray_phi=np.zeros(n_rays), ray_domega=np.zeros(n_rays))
        rc = self.lib.b200rt_make_grid_sph(self.precision, n_rb, n_sb, n_theta, n_phi, out["radial_boundaries"],
                                           szamethod, raymethod, out["sza_boundaries"], out["pts_radii"],
                                           out["pts_sza"], out["ray_theta"], out["ray_phi"], out["ray_domega"])
        if rc != 0:
            raise B200RTError(f"b200rt_make_grid_sph: status {rc}")
        return out

    def set_grid(self, g):
        n_rb, n_sb, n_rays = len(g["radial_boundaries"]), len(g["sza_boundaries"]), len(g["ray_theta"])
        self._ck(self.lib.b200rt_set_grid_sph(self.h, n_rb, n_sb, n_rays, g["radial_boundaries"], g["sza_boundaries"],
                                              g["pts_radii"], g["pts_sza"], g["ray_theta"], g["ray_phi"],
                                              g["ray_domega"]))
        self.n_rb, self.n_sb, self.n_rays = n_rb, n_sb, n_rays
        self.n_vox = (n_rb - 1) * (n_sb - 1)
        self.cap = 2 * n_rb + n_sb

    def make_grid_pp(self, n_rb, n_theta, rb):
        """plane_parallel_grid<n_rb, n_theta>::setup_voxels / setup_rays"""
        out = dict(radial_boundaries=np.ascontiguousarray(rb, dtype=np.float64), pts_radii=np.zeros(n_rb - 1),
                   ray_theta=np.zeros(n_theta), ray_domega=np.zeros(n_theta))
        rc = self.lib.b200rt_make_grid_pp(self.precision, n_rb, n_theta, out["radial_boundaries"], out["pts_radii"],
                                          out["ray_theta"], out["ray_domega"])
        if rc != 0:
            raise B200RTError(f"b200rt_make_grid_pp: status {rc}")
        return out

    def set_grid_pp(self, g):
        n_rb, n_rays = len(g["radial_boundaries"]), len(g["ray_theta"])
        self._ck(self.lib.b200rt_set_grid_pp(self.h, n_rb, n_rays, g["radial_boundaries"], g["pts_radii"],
                                             g["ray_theta"], g["ray_domega"]))
        self.n_rb, self.n_sb, self.n_rays = n_rb, 2, n_rays
        self.n_vox = n_rb - 1
        self.cap = 2 * n_rb + 2

    # ---- emissions
    def set_singlet(self, e, n_em, branching, T_ref, sigma_ref, g, tabs):
        """tabs: dict with the eight per-voxel tables of singlet_CFR (names as in the reference)"""
        names = ("T_ratio", "density", "dtau_species", "dtau_absorber",
                 "T_ratio_pt", "density_pt", "dtau_species_pt", "dtau_absorber_pt")
        arrs = [np.ascontiguousarray(tabs[k], dtype=np.float64) for k in names]
        self._ck(self.lib.b200rt_set_singlet(self.h, e, n_em, branching, T_ref, sigma_ref, g, *arrs))
        self.n_em = n_em

    # ---- source function
    def generate_S(self):
        self._ck(self.lib.b200rt_generate_S(self.h))

    def influence(self, v_begin=0, v_end=None):
        self._ck(self.lib.b200rt_influence(self.h, v_begin, self.n_vox if v_end is None else v_end))

    def solve(self):
        self._ck(self.lib.b200rt_solve(self.h))

    def last_step_count(self):
        n = C.c_longlong(0)
        self._ck(self.lib.b200rt_last_step_count(self.h, C.byref(n)))
        return n.value

    def solution(self, e, want_S=True):
        S = np.zeros(self.n_vox) if want_S else None
        S0, tsp, tab = np.zeros(self.n_vox), np.zeros(self.n_vox), np.zeros(self.n_vox)
        self._ck(self.lib.b200rt_get_solution(self.h, e, _ptr(S), _ptr(S0), _ptr(tsp), _ptr(tab)))
        return dict(S=S, S0=S0, tau_species_ss=tsp, tau_absorber_ss=tab)

    def influence_matrix(self, e, layout=ROW_MAJOR):
        K = np.zeros((self.n_vox, self.n_vox))
        self._ck(self.lib.b200rt_get_influence(self.h, e, layout, K))
        return K

    def set_sourcefn(self, e, S):
        self._ck(self.lib.b200rt_set_sourcefn(self.h, e, np.ascontiguousarray(S, dtype=np.float64)))

    def residual(self, e):
        r = C.c_double(0)
        self._ck(self.lib.b200rt_last_residual(self.h, e, C.byref(r)))
        return r.value

    def influence_dev(self, e):
        p = [_vp() for _ in range(4)]
        self._ck(self.lib.b200rt_influence_dev(self.h, e, *[C.byref(x) for x in p]))
        return [x.value for x in p]

    # ---- observations
    def los_from_MSO(self, locs, dirs):
        n = len(locs)
        out = [np.zeros(n) for _ in range(9)]
        rc = self.lib.b200rt_los_from_MSO(self.precision, n, np.ascontiguousarray(locs, dtype=np.float64),
                                          np.ascontiguousarray(dirs, dtype=np.float64), *out)
        if rc != 0:
            raise B200RTError(f"b200rt_los_from_MSO: status {rc}")
        return out

    def brightness(self, los, n_subsamples=10):
        """host buffers in, host buffers out: the reference-facing call (brightness_gpu)."""
        n = len(los[0])
        out = [np.zeros((self.n_em, n)) for _ in range(4)]
        self._ck(self.lib.b200rt_brightness(self.h, n, *los, n_subsamples, *[_ptr(o) for o in out]))
        self.n_los = n
        return dict(brightness=out[0], tau_species_final=out[1], tau_absorber_final=out[2], species_col_dens=out[3])

    def los_upload(self, los):
        self.n_los = len(los[0])
        self._ck(self.lib.b200rt_los_upload(self.h, self.n_los, *los))

    def brightness_resident(self, n_subsamples=10):
        self._ck(self.lib.b200rt_brightness_resident(self.h, n_subsamples))

    def los_download(self):
        out = [np.zeros((self.n_em, self.n_los)) for _ in range(4)]
        self._ck(self.lib.b200rt_los_download(self.h, *[_ptr(o) for o in out]))
        return dict(brightness=out[0], tau_species_final=out[1], tau_absorber_final=out[2], species_col_dens=out[3])

    # ---- interplanetary hydrogen background
    def iph_load_table(self, fname):
        self._ck(self.lib.b200rt_iph_load_table(self.h, os.fsencode(fname)))

    def iph_set_table(self, tab):
        """tab: dict with kmax, lmax, ninf, temp, alt_au, ang, dans, sot, so, sn, dinf_cm3 (file layout)"""
        f = lambda k: np.ascontiguousarray(tab[k], dtype=np.float32)
        self._ck(self.lib.b200rt_iph_set_table(self.h, int(tab["kmax"]), int(tab["lmax"]), int(tab["ninf"]),
                                               float(tab["temp"]), f("alt_au"), f("ang"), f("dans"), f("sot"), f("so"),
                                               f("sn"), f("dinf_cm3")))

    def iph_background(self, fs, pos, u, v, w, want_steps=False):
        n = len(u)
        fln = np.zeros(n, np.float32)
        steps = np.zeros(n, np.int32) if want_steps else None
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (u, v, w)]
        self._ck(self.lib.b200rt_iph_background(self.h, fs, pos[0], pos[1], pos[2], n, *a, fln, _ptr(steps)))
        return (fln, steps) if want_steps else fln

    def iph_model(self, g_lya, marspos, ra, dec):
        n = len(ra)
        out = np.zeros(n)
        self._ck(self.lib.b200rt_iph_model(self.h, g_lya, np.ascontiguousarray(marspos, dtype=np.float64), n,
                                           np.ascontiguousarray(ra, dtype=np.float64),
                                           np.ascontiguousarray(dec, dtype=np.float64), out))
        return out

    def iph_extinction(self, iph, tau_abs):
        iph = np.ascontiguousarray(iph, dtype=np.float64)
        tau = np.ascontiguousarray(tau_abs, dtype=np.float64)
        out = np.zeros_like(iph)
        rc = self.lib.b200rt_iph_extinction(iph.size, iph.ravel(), tau.ravel(), out.ravel())
        if rc != 0:
            raise B200RTError(f"b200rt_iph_extinction: status {rc}")
        return out

    # ---- traversal (parity surface)
    def traverse_voxel_rays(self, v0=0, v1=None):
        v1 = self.n_vox if v1 is None else v1
        nr = (v1 - v0) * self.n_rays
        cap = nr * self.cap
        ln, eb = np.zeros(nr, np.int32), np.zeros(nr, np.int32)
        ent, dist = np.zeros(cap, np.int32), np.zeros(cap)
        n = C.c_longlong(0)
        self._ck(self.lib.b200rt_traverse_voxel_rays(self.h, v0, v1, cap, ln, eb, ent, dist, C.byref(n)))
        return ln, eb, ent[:n.value].copy(), dist[:n.value].copy()

    def traverse_los(self):
        cap = self.n_los * self.cap
        ln, eb = np.zeros(self.n_los, np.int32), np.zeros(self.n_los, np.int32)
        ent, dist = np.zeros(cap, np.int32), np.zeros(cap)
        n = C.c_longlong(0)
        self._ck(self.lib.b200rt_traverse_los(self.h, cap, ln, eb, ent, dist, C.byref(n)))
        return ln, eb, ent[:n.value].copy(), dist[:n.value].copy()

    # ---- timing
    def kernel_ms(self, phase):
        ms, nl = C.c_float(0), C.c_int(0)
        self._ck(self.lib.b200rt_last_kernel_ms(self.h, phase, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def synchronize(self):
        self._ck(self.lib.b200rt_synchronize(self.h))

    def fp64_peaks(self):
        """-> (DFMA TFLOP/s, DMMA TFLOP/s) measured on this device"""
        a, b = C.c_double(0), C.c_double(0)
        self._ck(self.lib.b200rt_measure_fp64_peaks(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value


def define_singlet_tables(scn, e, precision=F64):
    """Host restatement of singlet_CFR::define (reference emission/singlet_CFR.hpp:419-492):
    the eight per-voxel tables from the atmosphere's voxel averages / point values.
    Evaluated in the arithmetic of `precision`, left to right as the reference does."""
    rt = np.float64 if precision == F64 else np.float32
    n_avg, n_pt, T_avg, T_pt, a_avg, a_pt = (scn.vox_in[k].astype(rt) for k in range(6))
    T_ref, sigma_ref = rt(scn.em_scalars[e][1]), rt(scn.em_scalars[e][2])
    sig = rt(scn.abs_sigma[e])
    Tr, Tr_pt = T_ref / T_avg, T_ref / T_pt
    tabs = dict(T_ratio=Tr, T_ratio_pt=Tr_pt, density=n_avg, density_pt=n_pt,
                dtau_species=n_avg * sigma_ref * np.sqrt(Tr), dtau_species_pt=n_pt * sigma_ref * np.sqrt(Tr_pt),
                dtau_absorber=a_avg * sig, dtau_absorber_pt=a_pt * sig)
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in tabs.items()}


class GpuModel:
    """Scenario-level convenience: one synthetic Scenario (synth.py) on one GPU, with the
    method vocabulary the parity tests use (grid / traverse / build_rows / solve / brightness)."""

    def __init__(self, scn, precision="f64", device=0):
        self.scn = scn
        self.prec = F64 if precision == "f64" else F32
        self.ctx = Context(device, self.prec)
        if getattr(scn, "pp", False):
            self.g = self.ctx.make_grid_pp(scn.n_rb, scn.n_theta, scn.rb)
            self.ctx.set_grid_pp(self.g)
        else:
            self.g = self.ctx.make_grid(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, scn.rb, scn.szamethod, scn.raymethod)
            self.ctx.set_grid(self.g)
        self.n_vox, self.n_rays = self.ctx.n_vox, self.ctx.n_rays
        for e in range(scn.n_em):
            b, T, s, g = (float(x) for x in scn.em_scalars[e])
            self.ctx.set_singlet(e, scn.n_em, b, T, s, g, define_singlet_tables(scn, e, self.prec))

    def grid(self):
        nphi = self.scn.n_phi
        g = dict(self.g)
        if getattr(self.scn, "pp", False):
            return g
        g["ray_theta"] = self.g["ray_theta"][::nphi].copy()
        g["ray_phi"] = self.g["ray_phi"][:nphi].copy()
        return g

    def traverse_voxel_rays(self, v0=0, v1=None):
        return self.ctx.traverse_voxel_rays(v0, v1)

    def traverse_los(self, locs, dirs):
        self.ctx.los_upload(self.ctx.los_from_MSO(locs, dirs))
        return self.ctx.traverse_los()

    def build_rows(self, v0=0, v1=None):
        self.ctx.influence(v0, v1)
        return self.ctx.kernel_ms(PH_INFLUENCE)[0] * 1e-3, self.ctx.last_step_count()

    def solve(self):
        self.ctx.solve()
        return [self.ctx.residual(e) for e in range(self.scn.n_em)]

    def K(self, e):
        return self.ctx.influence_matrix(e)

    def vectors(self, e, want_S=True):
        return self.ctx.solution(e, want_S)

    def set_sourcefn(self, e, S):
        self.ctx.set_sourcefn(e, S)

    def brightness(self, locs, dirs, n_subsamples=10):
        los = self.ctx.los_from_MSO(locs, dirs)
        r = self.ctx.brightness(los, n_subsamples)
        out = np.stack([r["brightness"], r["tau_species_final"], r["tau_absorber_final"], r["species_col_dens"]], axis=1)
        return self.ctx.kernel_ms(PH_BRIGHTNESS)[0] * 1e-3, out
