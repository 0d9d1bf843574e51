"""ctypes binding for oracle/librt_oracle_f{64,32}.so (rt_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Same method names as refbind.RefModel so tests can swap one for the other.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib_path(precision: str = "f64") -> str:
    return os.path.join(HERE, f"librt_oracle_{precision}.so")


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


_libs = {}


def _load(precision: str):
    if precision in _libs:
        return _libs[precision]
    if not os.path.exists(lib_path(precision)):
        build()
    lib = C.CDLL(lib_path(precision))
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.c_int] * 4 + [_dp, C.c_int, C.c_int]
    lib.oracle_create_pp.restype = C.c_void_p
    lib.oracle_create_pp.argtypes = [C.c_int, C.c_int, _dp]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_get_grid.argtypes = [C.c_void_p] + [_dp] * 6
    lib.oracle_define_singlet.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 5 + [_dp]
    lib.oracle_get_arrays.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.oracle_traverse_voxel_rays.restype = C.c_long
    lib.oracle_traverse_voxel_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _ip, _ip, _ip, _dp]
    lib.oracle_traverse_los.restype = C.c_long
    lib.oracle_traverse_los.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_long, _ip, _ip, _ip, _dp, _dp]
    lib.oracle_build_rows.restype = C.c_long
    lib.oracle_build_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.oracle_solve.restype = C.c_double
    lib.oracle_solve.argtypes = [C.c_void_p, C.c_int]
    lib.oracle_get_K.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.oracle_get_vectors.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    lib.oracle_set_sourcefn.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.oracle_brightness.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp]
    _libs[precision] = lib
    return lib


class OracleModel:
    def __init__(self, scn, precision: str = "f64"):
        self.lib = _load(precision)
        self.scn = scn
        self.n_vox = scn.n_vox
        self.n_rays = scn.n_rays
        if getattr(scn, "pp", False):      # plane_parallel_grid<n_rb, n_theta>
            self.h = self.lib.oracle_create_pp(scn.n_rb, scn.n_theta, np.ascontiguousarray(scn.rb))
        else:
            self.h = self.lib.oracle_create(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi,
                                            np.ascontiguousarray(scn.rb), scn.szamethod, scn.raymethod)
        for e in range(scn.n_em):
            b, T, s, g = (float(x) for x in scn.em_scalars[e])
            self.lib.oracle_define_singlet(self.h, e, b, T, s, g, float(scn.abs_sigma[e]),
                                           np.ascontiguousarray(scn.vox_in))

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def grid(self):
        s = self.scn
        out = dict(sza_boundaries=np.zeros(s.n_sb), pts_radii=np.zeros(s.n_rb - 1), pts_sza=np.zeros(s.n_sb - 1),
                   ray_theta=np.zeros(s.n_theta), ray_phi=np.zeros(s.n_phi), ray_domega=np.zeros(self.n_rays))
        self.lib.oracle_get_grid(self.h, out["sza_boundaries"], out["pts_radii"], out["pts_sza"],
                                 out["ray_theta"], out["ray_phi"], out["ray_domega"])
        return out

    ARRAY_NAMES = ("T_ratio", "T_ratio_pt", "density", "density_pt", "dtau_species", "dtau_species_pt",
                   "dtau_absorber", "dtau_absorber_pt", "abs", "abs_pt")

    def arrays(self, e: int):
        out = np.zeros((10, self.n_vox))
        self.lib.oracle_get_arrays(self.h, e, out)
        return dict(zip(self.ARRAY_NAMES, out))

    def traverse_voxel_rays(self, v0: int = 0, v1: int | None = None):
        v1 = self.n_vox if v1 is None else v1
        nr = (v1 - v0) * self.n_rays
        cap = nr * (2 * self.scn.n_rb + self.scn.n_sb)
        ln = np.zeros(nr, np.int32)
        eb = np.zeros(nr, np.int32)
        ent = np.zeros(cap, np.int32)
        dist = np.zeros(cap)
        n = self.lib.oracle_traverse_voxel_rays(self.h, v0, v1, cap, ln, eb, ent, dist)
        assert n >= 0
        return ln, eb, ent[:n].copy(), dist[:n].copy()

    def traverse_los(self, locs, dirs):
        n = len(locs)
        cap = n * (2 * self.scn.n_rb + self.scn.n_sb)
        ln = np.zeros(n, np.int32)
        eb = np.zeros(n, np.int32)
        ent = np.zeros(cap, np.int32)
        dist = np.zeros(cap)
        rs = np.zeros((n, 6))
        m = self.lib.oracle_traverse_los(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                                         np.ascontiguousarray(dirs, dtype=np.float64), cap, ln, eb, ent, dist, rs)
        assert m >= 0
        return ln, eb, ent[:m].copy(), dist[:m].copy(), rs

    def build_rows(self, v0=0, v1=None, stride=1):
        import time
        v1 = self.n_vox if v1 is None else v1
        t0 = time.perf_counter()
        ns = self.lib.oracle_build_rows(self.h, v0, v1, stride)
        return time.perf_counter() - t0, ns

    def solve(self):
        """-> list of relative residuals, one per emission"""
        return [self.lib.oracle_solve(self.h, e) for e in range(self.scn.n_em)]

    def K(self, e: int):
        out = np.zeros((self.n_vox, self.n_vox))
        self.lib.oracle_get_K(self.h, e, out)
        return out

    def vectors(self, e: int):
        S0, tsp, tab, S = (np.zeros(self.n_vox) for _ in range(4))
        self.lib.oracle_get_vectors(self.h, e, S0, tsp, tab, S)
        return dict(S0=S0, tau_species_ss=tsp, tau_absorber_ss=tab, S=S)

    def set_sourcefn(self, e: int, S):
        self.lib.oracle_set_sourcefn(self.h, e, np.ascontiguousarray(S, dtype=np.float64))

    def brightness(self, locs, dirs, n_subsamples: int = 10):
        import time
        n = len(locs)
        out = np.zeros((self.scn.n_em, 4, n))
        t0 = time.perf_counter()
        self.lib.oracle_brightness(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                                   np.ascontiguousarray(dirs, dtype=np.float64), n_subsamples, out)
        return time.perf_counter() - t0, out
