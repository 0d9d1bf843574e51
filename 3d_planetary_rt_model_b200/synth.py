"""Synthetic Chamberlain-exosphere inputs for the hot path (numpy, host side).

The reference builds its per-voxel tables from ``chamb_diff_1d`` (src/atm/*, Boost
gamma_p / odeint / B-splines), which SURVEY.md section 8 marks OUT OF SCOPE: parity is
evaluated at the *voxel-array boundary* -- the same arrays are handed to the
reference build, the oracle and the CUDA path.  This module only has to produce
physically sensible arrays of the right shape:

* temperature: Krasnopolsky profile (reference src/atm/temperature.cpp:23-48)
* exosphere: Chamberlain density without satellite particles
  (src/atm/chamberlain_exosphere.cpp:22-59) with P(3/2,x)=erf(sqrt x)-2 sqrt(x/pi) e^-x
* thermosphere: diffusive-equilibrium H in CO2, RK4 from the exobase down
  (src/atm/species_density_parameters.cpp:83-156, simplified)
* grids: radial boundaries `rmethod_altitude` / `rmethod_log_n_species`
  (src/grid/coordinate_generation.hpp:57-87, grid_spherical_azimuthally_symmetric.hpp:171-187)

It also generates line-of-sight sets: the reference's ``observation::fake`` image
(src/observation.hpp:173-208) and the seeded IUVS-like random set of SURVEY.md 8(d)
config 3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.special import erf

# physical constants, values as in reference src/constants.hpp:9-63
rMars = 3395e5
mMars = 0.1076 * 5.98e27
G = 6.67e-8
kB = 1.38e-16
clight = 3e10
mH = 1.673e-24
mCO2 = 44 * mH
line_f_coeff = 2.647e-2
aMars_typical = 1.41

lyman_alpha_lambda = 121.6e-7
lyman_alpha_f = 0.41641
lyman_alpha_cross_section_total = line_f_coeff * lyman_alpha_f
lyman_alpha_line_center_cross_section_coef = (
    lyman_alpha_cross_section_total / math.sqrt(2.0 * math.pi * kB / mH) * lyman_alpha_lambda)
CO2_lyman_alpha_absorption_cross_section = 6.3e-20
lyman_alpha_flux_Earth_typical = 4.5e15
lyman_alpha_flux_Mars_typical = (lyman_alpha_flux_Earth_typical / 1e4 * 1e8 * lyman_alpha_lambda
                                 * lyman_alpha_lambda / clight / aMars_typical / aMars_typical)
lyman_alpha_typical_g_factor = lyman_alpha_flux_Mars_typical * lyman_alpha_cross_section_total

lyman_beta_lambda = 102.6e-7
lyman_beta_f = 0.079142
lyman_beta_cross_section_total = line_f_coeff * lyman_beta_f
lyman_beta_branching_ratio = 0.8819
lyman_beta_line_center_cross_section_coef = (
    lyman_beta_cross_section_total / math.sqrt(2.0 * math.pi * kB / mH) * lyman_beta_lambda)
CO2_lyman_beta_absorption_cross_section = 3.53e-17
lyman_beta_flux_Earth_typical = lyman_alpha_flux_Earth_typical / 66.0
lyman_beta_flux_Mars_typical = (lyman_beta_flux_Earth_typical / 1e4 * 1e8 * lyman_beta_lambda
                                * lyman_beta_lambda / clight / aMars_typical / aMars_typical)
lyman_beta_typical_g_factor = lyman_beta_flux_Mars_typical * lyman_beta_cross_section_total


def _P32(x):
    """regularised lower incomplete gamma P(3/2, x)."""
    x = np.maximum(np.asarray(x, dtype=np.float64), 0.0)
    return erf(np.sqrt(x)) - 2.0 * np.sqrt(x / np.pi) * np.exp(-x)


@dataclass
class ChamberlainAtmosphere:
    """1-D (radial) H + CO2 atmosphere; all lengths in cm, densities cm^-3."""
    nH_exo: float = 5e5
    T_exo: float = 200.0
    nCO2_exo: float = 2e8
    rmin: float = rMars + 80e5
    rexo: float = rMars + 200e5
    n_species_min: float = 10.0
    rmax: float | None = None          # None => radius where n_H == n_species_min
    T_tropo: float = 125.0
    r_tropo: float = rMars + 90e5
    shape: float = 11.4
    m_species: float = mH              # 16*mH for the O I 102.6 nm scenarios
    _thermo_r: np.ndarray = field(init=False, repr=False, default=None)

    def __post_init__(self):
        self.lambdac = G * mMars * self.m_species / (kB * self.T_exo * self.rexo)
        veff = 0.5 * math.sqrt(2.0 * kB * self.T_exo / (self.m_species * math.pi)) * (1.0 + self.lambdac) * math.exp(-self.lambdac)
        self.escape_flux = self.nH_exo * veff
        if self.rmax is None:
            lo, hi = self.rexo, self.rexo * 1000.0
            for _ in range(200):
                mid = math.sqrt(lo * hi)
                if float(self._n_exo(mid)) > self.n_species_min:
                    lo = mid
                else:
                    hi = mid
            self.rmax = 0.5 * (lo + hi)
        self._integrate_thermosphere()

    # temperature ---------------------------------------------------------
    def Temp(self, r):
        r = np.asarray(r, dtype=np.float64)
        x = (r - self.r_tropo) * 1e-5
        sig = self.shape * self.T_exo
        T = np.where(x > 0, self.T_exo - (self.T_exo - self.T_tropo) * np.exp(-x * x / sig), self.T_tropo)
        return np.where(r > self.rexo, self.T_exo, T)

    def _Tprime(self, r):
        x = (r - self.r_tropo) * 1e-5
        sig = self.shape * self.T_exo
        if x <= 0:
            return 0.0
        T = self.T_exo - (self.T_exo - self.T_tropo) * math.exp(-x * x / sig)
        return (self.T_exo - T) * (2 * x / sig) * 1e-5

    # exosphere -----------------------------------------------------------
    def _n_exo(self, r):
        r = np.maximum(np.asarray(r, dtype=np.float64), self.rexo)
        lam = G * mMars * self.m_species / (kB * self.T_exo * r)
        psi = lam * lam / (lam + self.lambdac)
        frac = (1.0 + _P32(lam) - np.sqrt(np.maximum(1.0 - lam * lam / self.lambdac ** 2, 0.0))
                * np.exp(-psi) * (1.0 + _P32(lam - psi)))
        frac = frac / (1.0 + _P32(self.lambdac))
        return self.nH_exo * frac * np.exp(lam - self.lambdac)

    # thermosphere --------------------------------------------------------
    def _integrate_thermosphere(self, nsteps: int = 400):
        alpha = -0.25

        def deriv(r, y):
            lnCO2, lnH = y
            nCO2, nH = math.exp(lnCO2), math.exp(lnH)
            T = float(self.Temp(r)) if r <= self.rexo else self.T_exo
            Tp = self._Tprime(r)
            D = T ** 0.6 * 8.4e17 / nCO2
            K = 1.2e12 * math.sqrt(self.T_exo / nCO2)
            Hn_inv = G * mMars * mCO2 / (kB * T * r * r) + Tp / T
            HH_inv = G * mMars * self.m_species / (kB * T * r * r) + (1 + alpha) * Tp / T
            dCO2 = -Hn_inv
            dH = -(self.escape_flux * (self.rexo / r) ** 2 / nH + D * HH_inv + K * Hn_inv) / (D + K)
            return np.array([dCO2, dH])

        rs = np.linspace(self.rexo, self.rmin, nsteps)
        h = rs[1] - rs[0]
        y = np.array([math.log(self.nCO2_exo), math.log(self.nH_exo)])
        out = [y.copy()]
        for i in range(nsteps - 1):
            r = rs[i]
            k1 = deriv(r, y)
            k2 = deriv(r + 0.5 * h, y + 0.5 * h * k1)
            k3 = deriv(r + 0.5 * h, y + 0.5 * h * k2)
            k4 = deriv(r + h, y + h * k3)
            y = y + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            out.append(y.copy())
        out = np.array(out)
        self._thermo_r = rs[::-1].copy()
        self._thermo_lnCO2 = out[::-1, 0].copy()
        self._thermo_lnH = out[::-1, 1].copy()

    def n_species(self, r):
        r = np.asarray(r, dtype=np.float64)
        below = np.exp(np.interp(r, self._thermo_r, self._thermo_lnH))
        return np.where(r >= self.rexo, self._n_exo(r), below)

    def n_absorber(self, r):
        r = np.asarray(r, dtype=np.float64)
        below = np.exp(np.interp(r, self._thermo_r, self._thermo_lnCO2))
        # CO2 above the exobase: isothermal barometric fall-off, zero above rexo + 500 km
        lam = G * mMars * mCO2 / (kB * self.T_exo)
        above = self.nCO2_exo * np.exp(lam * (1.0 / np.maximum(r, self.rexo) - 1.0 / self.rexo))
        above = np.where(r > self.rexo + 500e5, 0.0, above)
        return np.where(r >= self.rexo, above, below)

    def r_from_n_species(self, n):
        lo, hi = self.rmin, self.rmax
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if float(self.n_species(mid)) > n:
                lo = mid
            else:
                hi = mid
        return 0.5 * (lo + hi)

    def sH_lya(self, T):
        return lyman_alpha_line_center_cross_section_coef / math.sqrt(T)

    def sH_lyb(self, T):
        return lyman_beta_line_center_cross_section_coef / math.sqrt(T)


# ------------------------------------------------------------------ grids
RMETHOD_ALTITUDE = 0
RMETHOD_LOG_N_SPECIES = 1
SZAMETHOD_UNIFORM = 0
SZAMETHOD_UNIFORM_COS = 1
RAYMETHOD_GAUSS = 0
RAYMETHOD_UNIFORM = 1


def radial_boundaries(atm: ChamberlainAtmosphere, n_rb: int, rmethod: int = RMETHOD_ALTITUDE) -> np.ndarray:
    """Radial boundaries; formulas of coordinate_generation.hpp:57-87 and
    grid_spherical_azimuthally_symmetric.hpp:178-187."""
    if rmethod == RMETHOD_ALTITUDE:
        nbelow = n_rb // 2
        logmax = math.log(atm.rmax - rMars)
        logmin = math.log(atm.rexo - rMars)
        logspace = (logmax - logmin) / float(n_rb - nbelow)
        linspace = (atm.rexo - atm.rmin) / float(nbelow - 1)
        rb = []
        for i in range(n_rb):
            if i < nbelow:
                rb.append(atm.rmin + i * linspace)
            else:
                rb.append(math.exp(logmin + (i - nbelow + 1) * logspace) + rMars)
        return np.array(rb, dtype=np.float64)
    lmax = math.log(float(atm.n_species(atm.rmin)))
    lmin = math.log(float(atm.n_species(atm.rmax)))
    step = (lmax - lmin) / (n_rb - 1.0)
    rb = [atm.r_from_n_species(math.exp(lmax - i * step)) for i in range(n_rb)]
    rb[0], rb[-1] = atm.rmin, atm.rmax
    return np.array(rb, dtype=np.float64)


def shell_average(f, r0: float, r1: float, npts: int = 48) -> float:
    """volume average of f(r) over the spherical shell [r0, r1] (Gauss-Legendre in log r)."""
    x, w = np.polynomial.legendre.leggauss(npts)
    l0, l1 = math.log(r0), math.log(r1)
    lr = 0.5 * (l1 + l0) + 0.5 * (l1 - l0) * x
    r = np.exp(lr)
    jac = r ** 3  # r^2 dr = r^3 dlnr
    return float(np.sum(w * f(r) * jac) / np.sum(w * jac))


def voxel_tables(atm: ChamberlainAtmosphere, rb: np.ndarray, n_sza_boundaries: int,
                 sza_T_contrast: float = 0.0) -> np.ndarray:
    """[6][n_vox] = n_avg, n_pt, T_avg, T_pt, nabs_avg, nabs_pt; voxel id = ir*(NSZA-1)+isza.

    sza_T_contrast != 0 adds a smooth day/night temperature variation so that tests
    exercise SZA-dependent tables (the kernels take fully per-voxel arrays).
    """
    n_r = len(rb) - 1
    n_s = n_sza_boundaries - 1
    out = np.zeros((6, n_r * n_s), dtype=np.float64)
    pts = np.sqrt(rb[:-1] * rb[1:])
    for i in range(n_r):
        navg = shell_average(atm.n_species, rb[i], rb[i + 1])
        Tavg = shell_average(atm.Temp, rb[i], rb[i + 1])
        aavg = shell_average(atm.n_absorber, rb[i], rb[i + 1])
        npt = float(atm.n_species(pts[i]))
        Tpt = float(atm.Temp(pts[i]))
        apt = float(atm.n_absorber(pts[i]))
        for j in range(n_s):
            f = 1.0 + sza_T_contrast * math.cos(math.pi * j / max(n_s - 1, 1))
            v = i * n_s + j
            out[:, v] = (navg, npt, Tavg * f, Tpt * f, aavg, apt)
    return out


@dataclass
class Scenario:
    """Everything the hot path needs for one (grid, atmosphere, emissions) case."""
    n_rb: int
    n_sb: int
    n_theta: int
    n_phi: int
    rb: np.ndarray                 # [n_rb] radial boundaries
    rexo: float
    szamethod: int
    raymethod: int
    em_scalars: np.ndarray         # [n_em][4] branching, T_ref, sigma_ref, g
    abs_sigma: np.ndarray          # [n_em]
    vox_in: np.ndarray             # [6][n_vox]
    pp: bool = False               # plane_parallel_grid<n_rb, n_theta> (n_sb = 2, n_phi = 1)

    @property
    def n_vox(self):
        return (self.n_rb - 1) * (self.n_sb - 1)

    @property
    def n_rays(self):
        return self.n_theta * self.n_phi

    @property
    def n_em(self):
        return len(self.abs_sigma)


def make_scenario(n_rb=40, n_sb=20, n_theta=7, n_phi=12, n_em=2, rmethod=RMETHOD_ALTITUDE,
                  nH_exo=5e5, T_exo=200.0, nCO2_exo=2e8, rmax=None, sza_T_contrast=0.0,
                  szamethod=SZAMETHOD_UNIFORM_COS, raymethod=RAYMETHOD_UNIFORM) -> Scenario:
    """Config-1-like scenario: H Ly alpha (+ Ly beta) singlet CFR."""
    atm = ChamberlainAtmosphere(nH_exo=nH_exo, T_exo=T_exo, nCO2_exo=nCO2_exo, rmax=rmax)
    rb = radial_boundaries(atm, n_rb, rmethod)
    vox = voxel_tables(atm, rb, n_sb, sza_T_contrast)
    em = [[1.0, T_exo, atm.sH_lya(T_exo), lyman_alpha_typical_g_factor],
          [lyman_beta_branching_ratio, T_exo, atm.sH_lyb(T_exo), lyman_beta_typical_g_factor]][:n_em]
    sig = [CO2_lyman_alpha_absorption_cross_section, CO2_lyman_beta_absorption_cross_section][:n_em]
    return Scenario(n_rb, n_sb, n_theta, n_phi, rb, atm.rexo, szamethod, raymethod,
                    np.array(em, dtype=np.float64), np.array(sig, dtype=np.float64), vox)


def make_scenario_pp(n_rb=40, n_theta=7, n_em=2, nH_exo=5e5, T_exo=200.0, nCO2_exo=2e8,
                     rmethod=RMETHOD_LOG_N_SPECIES) -> Scenario:
    """Plane-parallel scenario (reference grid/grid_plane_parallel.hpp; observation_fit uses <40, 7> with
    rmethod_log_n_species, observation_fit.cpp:28): shell averages are slab averages here
    (atmosphere_average_1d.cpp:141-157, the non-spherical branch)."""
    atm = ChamberlainAtmosphere(nH_exo=nH_exo, T_exo=T_exo, nCO2_exo=nCO2_exo)
    rb = radial_boundaries(atm, n_rb, rmethod)
    vox = voxel_tables(atm, rb, 2)
    em = [[1.0, T_exo, atm.sH_lya(T_exo), lyman_alpha_typical_g_factor],
          [lyman_beta_branching_ratio, T_exo, atm.sH_lyb(T_exo), lyman_beta_typical_g_factor]][:n_em]
    sig = [CO2_lyman_alpha_absorption_cross_section, CO2_lyman_beta_absorption_cross_section][:n_em]
    return Scenario(n_rb, 2, n_theta, 1, rb, atm.rexo, SZAMETHOD_UNIFORM_COS, RAYMETHOD_GAUSS,
                    np.array(em, dtype=np.float64), np.array(sig, dtype=np.float64), vox, pp=True)


# ------------------------------------------------------------------ multiplet emissions
MULT_O1026, MULT_H_LYMAN, MULT_H_SINGLET = 0, 1, 2      # O_1026_emission, H_lyman_multiplet, H_lyman_singlet
# (n_lines, n_multiplets, n_lower, n_upper, n_lambda): O_1026_tracker.hpp:12-15,188; H_multiplet_tracker.hpp:12-15,146
MULT_DIMS = {MULT_O1026: (6, 3, 3, 3, 21), MULT_H_LYMAN: (4, 2, 1, 4, 41), MULT_H_SINGLET: (2, 2, 1, 2, 41)}


@dataclass
class MultipletScenario:
    """One multiplet CFR emission (reference emission/multiplet_CFR_emission.hpp) on a spherical grid."""
    kind: int
    n_rb: int
    n_sb: int
    n_theta: int
    n_phi: int
    rb: np.ndarray
    rexo: float
    szamethod: int
    raymethod: int
    solar: np.ndarray              # [2] pumping flux [ph/cm2/s/Hz]: O I: (Lyman beta, -); H: (Lyman alpha, Lyman beta)
    vox_in: np.ndarray             # [6][n_vox] species avg/pt, temperature avg/pt, absorber avg/pt

    @property
    def n_vox(self):
        return (self.n_rb - 1) * (self.n_sb - 1)

    @property
    def n_rays(self):
        return self.n_theta * self.n_phi

    @property
    def dims(self):
        return MULT_DIMS[self.kind]


def make_multiplet_scenario(kind, n_rb=40, n_sb=20, n_theta=7, n_phi=12, sza_T_contrast=0.0,
                            szamethod=SZAMETHOD_UNIFORM_COS, raymethod=RAYMETHOD_UNIFORM) -> MultipletScenario:
    """BASELINE.json configs[4] (ii)/(iii).  O I: nO_exo = 2e7, T = 200 K, rmin = rMars + 100 km, solar Lyman beta
    1.69e-3 ph/cm2/s/Hz (generate_source_function.cpp:27,43-47,132; observation_fit.cpp:656-708).
    H: the config-1 atmosphere with the typical line-centre fluxes at Mars (constants.hpp:40-63)."""
    if kind == MULT_O1026:
        atm = ChamberlainAtmosphere(nH_exo=2e7, T_exo=200.0, nCO2_exo=2e8, rmin=rMars + 100e5, m_species=16 * mH)
        solar = np.array([1.69e-3, 0.0])
    else:
        atm = ChamberlainAtmosphere()
        solar = np.array([lyman_alpha_flux_Mars_typical, lyman_beta_flux_Mars_typical])
    rb = radial_boundaries(atm, n_rb, RMETHOD_ALTITUDE)
    vox = voxel_tables(atm, rb, n_sb, sza_T_contrast)
    return MultipletScenario(kind, n_rb, n_sb, n_theta, n_phi, rb, atm.rexo, szamethod, raymethod, solar, vox)


# ------------------------------------------------------------------ lines of sight
def _rodrigues(angle, ax):
    c, s = math.cos(angle), math.sin(angle)
    t = 1.0 - c
    x, y, z = ax
    return np.array([[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
                     [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                     [t * x * z - s * y, t * y * z + s * x, t * z * z + c]])


def fake_image(dist: float, angle_deg: float, nsamples: int, loc=(0.0, -1.0, 0.0)):
    """MSO positions / look directions of observation::fake (observation.hpp:173-208)."""
    loc = np.asarray(loc, dtype=np.float64)
    loc_norm = loc / math.sqrt(loc.dot(loc))
    pos = loc_norm * dist
    ang = math.pi / 180.0 * angle_deg
    dang = 2 * ang / (nsamples - 1)
    horiz = np.array([1.0, 0.0, 0.0])
    if pos[1] == 0.0 and pos[2] == 0.0:
        horiz = np.array([0.0, 1.0, 0.0])
    horiz = horiz - loc_norm * (horiz.dot(pos) / math.sqrt(pos.dot(pos)))
    horiz /= math.sqrt(horiz.dot(horiz))
    vert = np.cross(horiz, loc_norm)
    vert /= math.sqrt(vert.dot(vert))
    locs = np.tile(pos, (nsamples * nsamples, 1))
    dirs = np.empty_like(locs)
    for i in range(nsamples):
        Ri = _rodrigues(-ang + i * dang, horiz)
        for j in range(nsamples):
            R = Ri @ _rodrigues(-ang + j * dang, vert)
            dirs[i * nsamples + j] = -(R @ loc_norm)
    return locs, dirs


def random_los(n: int, seed: int = 20240607, r_lo: float = 1.05, r_hi: float = 2.85,
               cone_deg: float = 40.0):
    """Seeded IUVS-like LOS set: spacecraft radius uniform in
    [1.05, 2.85] rMars, position uniform on the sphere, look direction uniform in a cone
    of half-angle 40 deg about nadir.  Returns MSO (locs, dirs), float64 [n,3]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rad = rng.uniform(r_lo, r_hi, n) * rMars
    u = rng.uniform(-1.0, 1.0, n)
    ph = rng.uniform(0.0, 2 * math.pi, n)
    s = np.sqrt(1.0 - u * u)
    phat = np.stack([s * np.cos(ph), s * np.sin(ph), u], axis=1)
    locs = phat * rad[:, None]
    cmin = math.cos(math.radians(cone_deg))
    ct = rng.uniform(cmin, 1.0, n)
    st = np.sqrt(1.0 - ct * ct)
    az = rng.uniform(0.0, 2 * math.pi, n)
    # orthonormal frame about nadir (-phat)
    a = np.where(np.abs(phat[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    e1 = np.cross(phat, a)
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(phat, e1)
    dirs = (-phat) * ct[:, None] + e1 * (st * np.cos(az))[:, None] + e2 * (st * np.sin(az))[:, None]
    return np.ascontiguousarray(locs), np.ascontiguousarray(dirs)


# ------------------------------------------------------------------ interplanetary hydrogen (IPH) inputs
def make_iph_table(kmax: int = 59, lmax: int = 19, ninf: int = 5, temp: float = 8000.0):
    """A synthetic table with the layout of the reference's Quemerais model file
    (quemerais_IPH_model/fsm99td12v20t80: KMAX x LMAX nodes, 5 densities at infinity), for machines
    where that file is not available.  Physics-shaped, not physical: a hot-model-like ionisation
    cavity n/n_inf = exp(-r_c(theta)/r), optically thin source ~ n/r^2, primary and multiply
    scattered source functions attenuated / enhanced with the density at infinity."""
    alt = 0.2 * (551.6 / 0.2) ** (np.arange(kmax) / (kmax - 1.0)) ** 1.0
    alt = np.round(alt.astype(np.float64), 3)
    alt[0], alt[-1] = 0.2, 551.6
    ang = np.linspace(0.0, 180.0, lmax)
    th = np.radians(ang)[None, :]
    r = alt[:, None]
    r_c = 4.0 * (1.0 + 0.6 * (1.0 - np.cos(th)) / 2.0 * 3.0)          # cavity deeper downwind
    dans = np.exp(-r_c / r)
    dans[0, :] = 0.0
    dinf = 0.05 * (1 + np.arange(ninf))
    sot = dans / r ** 2
    so = np.empty((ninf, kmax, lmax))
    sn = np.empty((ninf, kmax, lmax))
    for i in range(ninf):
        tau = 0.8 * dinf[i] / 0.05 * np.sqrt(r)                        # grows outward and with density
        so[i] = dinf[i] * 1e6 * sot * np.exp(-0.3 * tau)               # like the real file: SO ~ SOT * n_inf [m^-3]
        sn[i] = so[i] * (1.0 + 0.25 * tau / (1.0 + 0.1 * tau))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(kmax=kmax, lmax=lmax, ninf=ninf, temp=temp, alt_au=f32(alt), ang=f32(ang), dans=f32(dans), sot=f32(sot),
                so=f32(so), sn=f32(sn), dinf_cm3=f32(dinf))


def random_sky(n: int, seed: int = 7):
    """n directions uniform on the sky -> (ra_deg, dec_deg)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    ra = rng.uniform(0.0, 360.0, n)
    dec = np.degrees(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    return ra, dec


MARS_ECLIPTIC_POS = (1.41, 0.3, 0.0)   # AU, SURVEY.md 8(d) config 3


def write_iph_table(tab, fname):
    """Write a table dict in the text layout of the reference's Quemerais file (the READ sequence of
    ipbackgroundCFR_fun.f:107-164: header lines, then 4 angle blocks of 5,5,5,4 columns per array)."""
    k, l, n = int(tab["kmax"]), int(tab["lmax"]), int(tab["ninf"])
    cols = [(0, 5), (5, 10), (10, 15), (15, 19)]

    def blocks(f, arr):
        for a, b in cols:
            f.write("       " + "".join(f"  {tab['ang'][j]:4.0f}.  " for j in range(a, b)) + "\n")
            for i in range(k):
                f.write(f"{tab['alt_au'][i]:9.3f}" + "".join(f" {arr[i, j]:13.6E}" for j in range(a, b)) + "\n")

    with open(fname, "w") as f:
        f.write(f"  {k}  {l}  {n}\n")
        for ii in range(n):
            f.write(f"   20.00  254.00   7.50 {tab['temp']:5.0f}.  0.99 0.120E+07 0.0 {tab['dinf_cm3'][ii]:9.3f}\n")
            if ii == 0:
                blocks(f, np.asarray(tab["dans"]).reshape(k, l))
                blocks(f, np.asarray(tab["sot"]).reshape(k, l))
            blocks(f, np.asarray(tab["so"]).reshape(n, k, l)[ii])
            blocks(f, np.asarray(tab["sn"]).reshape(n, k, l)[ii])
