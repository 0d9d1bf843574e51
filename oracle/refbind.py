"""ctypes binding for oracle/_ref/libref_f{64,32}.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own hot-path source built in place (oracle/Makefile,
oracle/ref_harness.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
reference / cpu_baseline legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def lib_path(precision: str = "f64", variant: str = "") -> str:
    """variant "b200": the same harness built with -DRT_B200, i.e. the reference's RT_grid with its *_gpu members
    bound to libb200rt.so by integration/RT_b200.hpp (oracle/Makefile ref_b200)"""
    return os.path.join(HERE, "_ref", f"libref_{variant + '_' if variant else ''}{precision}.so")


def available(precision: str = "f64", variant: str = "") -> bool:
    return os.path.exists(lib_path(precision, variant))


_libs = {}


def _load(precision: str, variant: str = ""):
    key = precision + variant
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(lib_path(precision, variant))
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int] * 5
    lib.ref_create_pp.restype = C.c_void_p
    lib.ref_create_pp.argtypes = [C.c_int] * 3
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_n_voxels.argtypes = [C.c_void_p]
    lib.ref_n_rays.argtypes = [C.c_void_p]
    lib.ref_setup.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    lib.ref_get_grid.argtypes = [C.c_void_p] + [_dp] * 7
    lib.ref_get_arrays.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_traverse_voxel_rays.restype = C.c_long
    lib.ref_traverse_voxel_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_long, _ip, _ip, _ip, _dp]
    lib.ref_traverse_los.restype = C.c_long
    lib.ref_traverse_los.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_long, _ip, _ip, _ip, _dp, _dp]
    lib.ref_generate_S.restype = C.c_double
    lib.ref_generate_S.argtypes = [C.c_void_p]
    lib.ref_build_rows.restype = C.c_double
    lib.ref_build_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
    lib.ref_solve.restype = C.c_double
    lib.ref_solve.argtypes = [C.c_void_p]
    lib.ref_get_K.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_get_vectors.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    lib.ref_set_sourcefn.argtypes = [C.c_void_p, C.c_int, _dp]
    lib.ref_brightness.restype = C.c_double
    lib.ref_brightness.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp]
    lib.ref_omp_threads.restype = C.c_int
    if hasattr(lib, "ref_save_S"):
        lib.ref_save_S.argtypes = [C.c_void_p, C.c_char_p]
        lib.ref_save_influence.argtypes = [C.c_void_p, C.c_char_p]
    if hasattr(lib, "ref_set_omp_threads"):
        lib.ref_set_omp_threads.argtypes = [C.c_int]
    lib.ref_real_bytes.restype = C.c_int
    if hasattr(lib, "ref_generate_S_gpu"):
        lib.ref_generate_S_gpu.argtypes = [C.c_void_p]
        lib.ref_influence_to_host.argtypes = [C.c_void_p]
        lib.ref_brightness_gpu.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp]
    _libs[key] = lib
    return lib


class RefModel:
    """One reference RT_grid<singlet_CFR, n_em, spherical_azimuthally_symmetric_grid<...>>."""

    def __init__(self, scn, precision: str = "f64", rmethod_inject: bool = True, variant: str = ""):
        self.lib = _load(precision, variant)
        self.scn = scn
        if getattr(scn, "pp", False):      # plane_parallel_grid<n_rb, n_theta>
            self.h = self.lib.ref_create_pp(scn.n_rb, scn.n_theta, scn.n_em)
        else:
            self.h = self.lib.ref_create(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, scn.n_em)
        if not self.h:
            raise ValueError(f"grid shape {(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi)} x{scn.n_em} "
                             "is not instantiated in oracle/ref_harness.cpp")
        self.n_vox = self.lib.ref_n_voxels(self.h)
        self.n_rays = self.lib.ref_n_rays(self.h)
        rc = self.lib.ref_setup(self.h, np.ascontiguousarray(scn.rb), float(scn.rexo),
                                1 if rmethod_inject else 0, scn.szamethod, scn.raymethod, scn.n_em,
                                np.ascontiguousarray(scn.em_scalars), np.ascontiguousarray(scn.abs_sigma),
                                np.ascontiguousarray(scn.vox_in))
        if rc != 0:
            raise RuntimeError(f"ref_setup failed: {rc}")

    def __del__(self):
        try:
            if self.h:
                self.lib.ref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def grid(self):
        s = self.scn
        out = dict(sza_boundaries=np.zeros(s.n_sb), pts_radii=np.zeros(s.n_rb - 1), pts_sza=np.zeros(s.n_sb - 1),
                   ray_theta=np.zeros(s.n_theta), ray_phi=np.zeros(s.n_phi), ray_domega=np.zeros(self.n_rays),
                   radial_boundaries=np.zeros(s.n_rb))
        self.lib.ref_get_grid(self.h, out["sza_boundaries"], out["pts_radii"], out["pts_sza"], out["ray_theta"],
                              out["ray_phi"], out["ray_domega"], out["radial_boundaries"])
        return out

    ARRAY_NAMES = ("T_ratio", "T_ratio_pt", "density", "density_pt", "dtau_species", "dtau_species_pt",
                   "dtau_absorber", "dtau_absorber_pt", "abs", "abs_pt")

    def arrays(self, e: int):
        out = np.zeros((10, self.n_vox))
        self.lib.ref_get_arrays(self.h, e, out)
        return dict(zip(self.ARRAY_NAMES, out))

    def traverse_voxel_rays(self, v0: int = 0, v1: int | None = None):
        v1 = self.n_vox if v1 is None else v1
        nr = (v1 - v0) * self.n_rays
        cap = nr * (2 * self.scn.n_rb + self.scn.n_sb)
        ln = np.zeros(nr, np.int32)
        eb = np.zeros(nr, np.int32)
        ent = np.zeros(cap, np.int32)
        dist = np.zeros(cap)
        n = self.lib.ref_traverse_voxel_rays(self.h, v0, v1, cap, ln, eb, ent, dist)
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
        m = self.lib.ref_traverse_los(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                                      np.ascontiguousarray(dirs, dtype=np.float64), cap, ln, eb, ent, dist, rs)
        assert m >= 0
        return ln, eb, ent[:m].copy(), dist[:m].copy(), rs

    def generate_S(self) -> float:
        return self.lib.ref_generate_S(self.h)

    def build_rows(self, v0=0, v1=None, stride=1):
        v1 = self.n_vox if v1 is None else v1
        ns = C.c_long(0)
        t = self.lib.ref_build_rows(self.h, v0, v1, stride, C.byref(ns))
        return t, ns.value

    def solve(self) -> float:
        return self.lib.ref_solve(self.h)

    def K(self, e: int):
        out = np.zeros((self.n_vox, self.n_vox))
        self.lib.ref_get_K(self.h, e, out)
        return out

    def vectors(self, e: int):
        S0, tsp, tab, S = (np.zeros(self.n_vox) for _ in range(4))
        self.lib.ref_get_vectors(self.h, e, S0, tsp, tab, S)
        return dict(S0=S0, tau_species_ss=tsp, tau_absorber_ss=tab, S=S)

    def set_sourcefn(self, e: int, S):
        self.lib.ref_set_sourcefn(self.h, e, np.ascontiguousarray(S, dtype=np.float64))

    def brightness(self, locs, dirs, n_subsamples: int = 10):
        """-> (seconds, out[n_em][4][n]) rows: brightness, tau_species, tau_absorber, col_dens"""
        n = len(locs)
        out = np.zeros((self.scn.n_em, 4, n))
        t = self.lib.ref_brightness(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                                    np.ascontiguousarray(dirs, dtype=np.float64), n_subsamples, out)
        return t, out

    # ---- the reference's RT_grid::*_gpu members, bound to libb200rt.so by integration/RT_b200.hpp (variant "b200")
    def generate_S_gpu(self) -> None:
        rc = self.lib.ref_generate_S_gpu(self.h)
        if rc != 0:
            raise RuntimeError(f"RT_grid::generate_S_gpu failed ({rc}): built without -DRT_B200, or no GPU")

    def influence_to_host(self) -> None:
        if self.lib.ref_influence_to_host(self.h) != 0:
            raise RuntimeError("RT_grid::emissions_influence_to_host failed")

    def brightness_gpu(self, locs, dirs, n_subsamples: int = 10):
        n = len(locs)
        out = np.zeros((self.scn.n_em, 4, n))
        rc = self.lib.ref_brightness_gpu(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                                         np.ascontiguousarray(dirs, dtype=np.float64), n_subsamples, out)
        if rc != 0:
            raise RuntimeError(f"RT_grid::brightness_gpu failed ({rc})")
        return out

    def save_S(self, fname: str) -> None:
        """RT_grid::save_S (RT_grid.hpp:228-230): the reference's own ASCII writer"""
        if self.lib.ref_save_S(self.h, os.fsencode(fname)) != 0:
            raise RuntimeError("save_S not available on this reference model")

    def save_influence(self, fname: str) -> None:
        """RT_grid::save_influence (RT_grid.hpp:221-227)"""
        if self.lib.ref_save_influence(self.h, os.fsencode(fname)) != 0:
            raise RuntimeError("save_influence not available on this reference model")

    def omp_threads(self) -> int:
        return self.lib.ref_omp_threads()

    def use_all_cores(self) -> int:
        """pin the OpenMP thread count to the cores this process may run on (torchrun exports OMP_NUM_THREADS=1,
        which would otherwise time the reference on one core) -> the thread count in effect"""
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        self.lib.ref_set_omp_threads(n)
        return self.lib.ref_omp_threads()
