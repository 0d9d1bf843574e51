"""ctypes bindings of the multiplet-emission checkers.  TEST INFRASTRUCTURE ONLY.

OracleMultiplet -> oracle/librt_oracle_f{64,32}.so (multiplet_oracle.inc.c, our restatement)
RefMultiplet    -> oracle/_ref/libref_mult_f{64,32}.so (the reference's own multiplet source, built in place)
Both expose the same vocabulary so that tests can swap them.
"""
import ctypes as C
import os

import numpy as np

from . import oraclebind

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def ref_lib_path(precision):
    return os.path.join(HERE, "_ref", f"libref_mult_{precision}.so")


def ref_available(precision="f64"):
    return os.path.exists(ref_lib_path(precision))


_libs = {}


def _load(which, precision):
    key = (which, precision)
    if key in _libs:
        return _libs[key]
    if which == "oracle":
        oraclebind._load(precision)          # builds if needed
        lib = C.CDLL(oraclebind.lib_path(precision))
        pre = "oracle_multiplet_"
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_int] * 4 + [_dp, C.c_int, C.c_int]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_define_multiplet.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        lib.oracle_multiplet_build_rows.restype = C.c_long
        lib.oracle_multiplet_build_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.oracle_multiplet_brightness.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp]
    else:
        lib = C.CDLL(ref_lib_path(precision))
        pre = "refm_"
        lib.refm_create.restype = C.c_void_p
        lib.refm_create.argtypes = [C.c_int] * 5
        lib.refm_destroy.argtypes = [C.c_void_p]
        lib.refm_setup.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, C.c_int, _dp, _dp]
        lib.refm_build_rows.restype = C.c_double
        lib.refm_build_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
        lib.refm_generate_S.restype = C.c_double
        lib.refm_generate_S.argtypes = [C.c_void_p]
        lib.refm_brightness.restype = C.c_double
        lib.refm_brightness.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int, _dp]
    g = lambda n: getattr(lib, pre + n)
    g("dims").argtypes = [C.c_void_p, _ip]
    g("constants").argtypes = [C.c_void_p, _ip, _dp]
    g("lineshape").restype = C.c_double
    g("lineshape").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
    g("get_arrays").argtypes = [C.c_void_p, _dp]
    g("solve").restype = C.c_double
    g("solve").argtypes = [C.c_void_p]
    g("get_K").argtypes = [C.c_void_p, _dp]
    g("get_vectors").argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    g("set_sourcefn").argtypes = [C.c_void_p, _dp]
    _libs[key] = (lib, pre)
    return _libs[key]


class _Common:
    def _f(self, name):
        return getattr(self.lib, self.pre + name)

    def _dims(self):
        o = np.zeros(7, np.int32)
        self._f("dims")(self.h, o)
        (self.n_vox, self.n_rays, self.n_lines, self.n_mult, self.n_lower, self.n_upper, self.n_lambda) = (int(x) for x in o)
        self.n_el = self.n_vox * self.n_upper

    def constants(self):
        """dict of the tracker constants: index arrays [n_lines] and the seven per-line Real arrays"""
        i = np.zeros(3 * self.n_lines, np.int32)
        d = np.zeros(7 * self.n_lines)
        self._f("constants")(self.h, i, d)
        i, d = i.reshape(3, -1), d.reshape(7, -1)
        return dict(multiplet_index=i[0], lower_level_index=i[1], upper_level_index=i[2], line_sigma_total=d[0],
                    line_A=d[1], absorber_xsec=d[2], upper_state_decay_rate=d[3][:self.n_upper], offset=d[4],
                    norm=d[5], weight=d[6])

    def lineshape(self, line, i_lambda, T):
        return self._f("lineshape")(self.h, line, i_lambda, float(T))

    def arrays(self):
        out = np.zeros((2 * self.n_lower + 4, self.n_vox))
        self._f("get_arrays")(self.h, out)
        d = {}
        for l in range(self.n_lower):
            d[f"species_density_{l}"], d[f"species_density_pt_{l}"] = out[2 * l], out[2 * l + 1]
        o = out[2 * self.n_lower:]
        d.update(species_T=o[0], species_T_pt=o[1], absorber_density=o[2], absorber_density_pt=o[3])
        return d

    def solve(self):
        return self._f("solve")(self.h)

    def K(self):
        out = np.zeros((self.n_el, self.n_el))
        self._f("get_K")(self.h, out)
        return out

    def vectors(self, want_S=True):
        S0, S = np.zeros(self.n_el), np.zeros(self.n_el)
        tsp, tab = np.zeros(self.n_vox * self.n_lines), np.zeros(self.n_vox * self.n_lines)
        self._f("get_vectors")(self.h, S0, tsp, tab, S)
        return dict(S0=S0, tau_species_ss=tsp, tau_absorber_ss=tab, S=S)

    def set_sourcefn(self, S):
        self._f("set_sourcefn")(self.h, np.ascontiguousarray(S, dtype=np.float64))

    def brightness(self, locs, dirs, n_subsamples=10):
        """-> dict of [n_lines][n] brightness, tau_species_final, tau_absorber_final and [n_lower][n] species_col_dens"""
        n = len(locs)
        out = np.zeros((3 * self.n_lines + self.n_lower, n))
        self._f("brightness")(self.h, n, np.ascontiguousarray(locs, dtype=np.float64),
                              np.ascontiguousarray(dirs, dtype=np.float64), n_subsamples, out)
        NL = self.n_lines
        return dict(brightness=out[:NL], tau_species_final=out[NL:2 * NL], tau_absorber_final=out[2 * NL:3 * NL],
                    species_col_dens=out[3 * NL:])


class OracleMultiplet(_Common):
    def __init__(self, scn, precision="f64"):
        self.lib, self.pre = _load("oracle", precision)
        self.scn = scn
        self.h = self.lib.oracle_create(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi, np.ascontiguousarray(scn.rb),
                                        scn.szamethod, scn.raymethod)
        self.lib.oracle_define_multiplet(self.h, scn.kind, np.ascontiguousarray(scn.solar, dtype=np.float64),
                                         np.ascontiguousarray(scn.vox_in))
        self._dims()

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build_rows(self, v0=0, v1=None, stride=1):
        import time
        t0 = time.perf_counter()
        ns = self.lib.oracle_multiplet_build_rows(self.h, v0, self.n_vox if v1 is None else v1, stride)
        return time.perf_counter() - t0, ns


class RefMultiplet(_Common):
    def __init__(self, scn, precision="f64"):
        self.lib, self.pre = _load("ref", precision)
        self.scn = scn
        self.h = self.lib.refm_create(scn.kind, scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi)
        if not self.h:
            raise ValueError(f"grid shape {(scn.n_rb, scn.n_sb, scn.n_theta, scn.n_phi)} kind {scn.kind} "
                             "is not instantiated in oracle/ref_harness_multiplet.cpp")
        self.lib.refm_setup(self.h, np.ascontiguousarray(scn.rb), float(scn.rexo), scn.szamethod, scn.raymethod,
                            np.ascontiguousarray(scn.solar, dtype=np.float64), np.ascontiguousarray(scn.vox_in))
        self._dims()

    def __del__(self):
        try:
            if self.h:
                self.lib.refm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def build_rows(self, v0=0, v1=None, stride=1):
        ns = C.c_long(0)
        t = self.lib.refm_build_rows(self.h, v0, self.n_vox if v1 is None else v1, stride, C.byref(ns))
        return t, ns.value

    def generate_S(self):
        return self.lib.refm_generate_S(self.h)
