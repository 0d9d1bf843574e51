// atmosphere.cpp -- see atmosphere.hpp
#include "atmosphere.hpp"
#include <algorithm>
#include <cmath>
#include <fstream>
#include <stdexcept>

namespace b200rt_host {

namespace {
inline double P32(double x) {   // regularised lower incomplete gamma P(3/2, x)
  x = std::max(x, 0.0);
  return std::erf(std::sqrt(x)) - 2.0 * std::sqrt(x / pi) * std::exp(-x);
}
inline double interp(double x, const std::vector<double> &xs, const std::vector<double> &ys) {   // numpy.interp
  if (x <= xs.front()) return ys.front();
  if (x >= xs.back()) return ys.back();
  const size_t j = std::upper_bound(xs.begin(), xs.end(), x) - xs.begin();   // xs[j-1] <= x < xs[j]
  const double slope = (ys[j] - ys[j - 1]) / (xs[j] - xs[j - 1]);
  return slope * (x - xs[j - 1]) + ys[j - 1];
}
const std::vector<double> &glx(int which) {
  static std::vector<double> x, w;
  if (x.empty()) gauss_legendre(48, x, w);
  return which ? w : x;
}
} // namespace

void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = std::cos(pi * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; it++) {
      double p1 = 1.0, p2 = 0.0;
      for (int j = 0; j < n; j++) {
        const double p3 = p2;
        p2 = p1;
        p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1.0);
      }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      const double z1 = z;
      z = z1 - p1 / pp;
      if (std::fabs(z - z1) < 1e-15) break;
    }
    x[i] = -z;
    x[n - 1 - i] = z;
    w[i] = 2.0 / ((1.0 - z * z) * pp * pp);
    w[n - 1 - i] = w[i];
  }
}

// ---------------------------------------------------------------- temperature
krasnopolsky_temperature::krasnopolsky_temperature(double T_exoo, double T_tropoo, double r_tropoo, double shape_parameterr,
                                                   bool shape_parameter_Texo)
    : T_exo(T_exoo), T_tropo(T_tropoo), r_tropo(r_tropoo),
      shape_parameter(shape_parameter_Texo ? shape_parameterr * T_exoo : shape_parameterr * shape_parameterr) {}

double krasnopolsky_temperature::T(double r) const {
  const double x = (r - r_tropo) * 1e-5;
  return (x > 0) ? T_exo - (T_exo - T_tropo) * std::exp(-x * x / shape_parameter) : T_tropo;
}

double krasnopolsky_temperature::Tprime(double r) const {
  const double x = (r - r_tropo) * 1e-5;
  if (x <= 0) return 0.0;
  return (T_exo - T(r)) * (2 * x / shape_parameter) * 1e-5;
}

// ---------------------------------------------------------------- the common part
Real atmosphere_model::r_from_n_species(Real n) const {
  Real lo = rmin, hi = rmax;
  for (int i = 0; i < 200; i++) {
    const Real mid = 0.5 * (lo + hi);
    if (n_species(mid) > n) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

std::vector<Real> atmosphere_model::radial_boundaries(int n_rb, int rmethod) const {
  std::vector<Real> rb(n_rb);
  if (rmethod == 0) {
    const int nbelow = n_rb / 2;
    const Real logmax = std::log(rmax - rMars), logmin = std::log(rexo - rMars);
    const Real logspace = (logmax - logmin) / Real(n_rb - nbelow);
    const Real linspace = (rexo - rmin) / Real(nbelow - 1);
    for (int i = 0; i < n_rb; i++)
      rb[i] = (i < nbelow) ? rmin + i * linspace : std::exp(logmin + (i - nbelow + 1) * logspace) + rMars;
    return rb;
  }
  const Real lmax = std::log(n_species(rmin)), lmin = std::log(n_species(rmax));
  const Real step = (lmax - lmin) / (n_rb - 1.0);
  for (int i = 0; i < n_rb; i++) rb[i] = r_from_n_species(std::exp(lmax - i * step));
  rb[0] = rmin;
  rb[n_rb - 1] = rmax;
  return rb;
}

template <class F>
Real atmosphere_model::shell_average(F f, Real r0, Real r1) const {
  const std::vector<double> &gx = glx(0), &gw = glx(1);
  Real num = 0, den = 0;
  if (spherical) {
    const Real l0 = std::log(r0), l1 = std::log(r1);
    for (size_t k = 0; k < gx.size(); k++) {
      const Real r = std::exp(0.5 * (l1 + l0) + 0.5 * (l1 - l0) * gx[k]);
      const Real jac = r * r * r;   // r^2 dr = r^3 dln r
      num += gw[k] * f(r) * jac;
      den += gw[k] * jac;
    }
  } else {
    for (size_t k = 0; k < gx.size(); k++) {
      const Real r = 0.5 * (r1 + r0) + 0.5 * (r1 - r0) * gx[k];
      num += gw[k] * f(r);
      den += gw[k];
    }
  }
  return num / den;
}

void atmosphere_model::voxel_values(Real r0, Real r1, Real, Real, Real pt_r, Real, Real (&out)[6]) const {
  out[0] = shell_average([&](Real r) { return n_species(r); }, r0, r1);
  out[1] = n_species(pt_r);
  if (temp_dependent_sH) {
    out[2] = shell_average([&](Real r) { return Temp(r); }, r0, r1);
    out[3] = Temp(pt_r);
  } else {
    out[2] = out[3] = constant_temp_sH;
  }
  out[4] = shell_average([&](Real r) { return n_absorber(r); }, r0, r1);
  out[5] = n_absorber(pt_r);
}

void atmosphere_model::voxel_tables(const std::vector<Real> &rb, const std::vector<Real> &pts_r, const std::vector<Real> &sb,
                                    const std::vector<Real> &pts_s, std::vector<Real> (&out)[6]) const {
  const int n_r = (int) rb.size() - 1, n_s = (int) sb.size() - 1;
  for (auto &v : out) v.assign((size_t) n_r * n_s, 0.0);
  const bool per_sza = sza_dependent();
  for (int i = 0; i < n_r; i++)
    for (int j = 0; j < n_s; j++) {
      Real vals[6];
      if (per_sza || j == 0) voxel_values(rb[i], rb[i + 1], sb[j], sb[j + 1], pts_r[i], pts_s[j], vals);
      else for (int q = 0; q < 6; q++) vals[q] = out[q][(size_t) i * n_s];
      for (int q = 0; q < 6; q++) out[q][(size_t) i * n_s + j] = vals[q];
    }
}

void atmosphere_model::voxel_tables(const std::vector<Real> &rb, int n_sb, std::vector<Real> (&out)[6]) const {
  const int n_r = (int) rb.size() - 1, n_s = n_sb - 1;
  for (auto &v : out) v.assign((size_t) n_r * n_s, 0.0);
  for (int i = 0; i < n_r; i++) {
    Real vals[6];
    atmosphere_model::voxel_values(rb[i], rb[i + 1], 0, pi, std::sqrt(rb[i] * rb[i + 1]), 0, vals);
    for (int j = 0; j < n_s; j++)
      for (int q = 0; q < 6; q++) out[q][(size_t) i * n_s + j] = vals[q];
  }
}

void atmosphere_model::save(const std::string &fname) const {
  std::ofstream file(fname.c_str());
  if (!file.is_open()) return;
  file << "b200rt atmosphere: rmin = " << rmin << " cm, rexo = " << rexo << " cm, rmax = " << rmax << " cm\n";
  file << "r [cm]   n_species [cm-3]   n_absorber [cm-3]   T [K]\n";
  const int n = 200;
  for (int i = 0; i < n; i++) {
    const Real r = rMars + (rmin - rMars) * std::pow((rmax - rMars) / (rmin - rMars), i / Real(n - 1));
    file << r << " " << n_species(r) << " " << n_absorber(r) << " " << Temp(r) << "\n";
  }
}

// ---------------------------------------------------------------- chamb_diff_1d
chamb_diff_1d::chamb_diff_1d(Real nHexo, Real nCO2exo, Real Texo, Real rmaxx)
    : nH_exo(nHexo), T_exo(Texo), nCO2_exo(nCO2exo), rmindiffusion(rMars + 80e5), temp(Texo),
      species(species_density_parameters::hydrogen()) {
  rmin = rMars + 80e5;
  rexo = rexo_typical;
  if (rmaxx > 0) { rmax = rmaxx; setup(rmaxx, nCO2exo, -1); }
  else setup(n_species_min, nCO2exo, method_nspmin_nCO2exo);
}

chamb_diff_1d::chamb_diff_1d(Real rminn, Real rexoo, Real rmaxx_or_nspmin, Real rmindiffusionn, Real n_species_exoo,
                             Real nCO2rmin_or_nCO2exoo, const krasnopolsky_temperature &tempp,
                             const species_density_parameters &sp, int method)
    : nH_exo(n_species_exoo), T_exo(tempp.T_exo), nCO2_exo(0), rmindiffusion(rmindiffusionn), temp(tempp), species(sp) {
  rmin = rminn;
  rexo = rexoo;
  setup(rmaxx_or_nspmin, nCO2rmin_or_nCO2exoo, method);
}

// method -1: rmax given, CO2 given at the exobase (setup_rmax_nCO2exo, thermosphere_exosphere.cpp:94-108)
void chamb_diff_1d::setup(Real rmaxx_or_nspmin, Real nCO2rmin_or_exo, int method) {
  lambdac = G * mMars * species.mass / (kB * T_exo * rexo);
  const Real veff = 0.5 * std::sqrt(2.0 * kB * T_exo / (species.mass * pi)) * (1.0 + lambdac) * std::exp(-lambdac);
  escape_flux = nH_exo * veff;
  if (method == method_nspmin_nCO2exo) {        // rmax = where the exosphere falls to n_species_min (:57-75)
    n_species_min = rmaxx_or_nspmin;
    Real lo = rexo, hi = rexo * 1000.0;
    for (int i = 0; i < 200; i++) {
      const Real mid = std::sqrt(lo * hi);
      if (n_exo(mid, nH_exo, species.mass) > n_species_min) lo = mid; else hi = mid;
    }
    rmax = 0.5 * (lo + hi);
    nCO2_exo = nCO2rmin_or_exo;
  } else if (method == method_rmax_nCO2rmin) {   // CO2 given at rmin: integrate it up to the exobase (:77-92)
    rmax = rmaxx_or_nspmin;
    nCO2_exo = CO2_exobase_from_rmin(nCO2rmin_or_exo);
  } else {
    rmax = rmaxx_or_nspmin;
    nCO2_exo = nCO2rmin_or_exo;
  }
  integrate_thermosphere();
  n_species_rmindiffusion = std::exp(interp(rmindiffusion, thermo_r, thermo_lnH));
  nCO2_rmindiffusion = std::exp(interp(rmindiffusion, thermo_r, thermo_lnCO2));
}

Real chamb_diff_1d::Temp(Real r) const { return (r > rexo) ? T_exo : temp.T(r); }

Real chamb_diff_1d::n_exo(Real r, Real n0, Real mass) const {
  r = std::max(r, rexo);
  const Real lc = G * mMars * mass / (kB * T_exo * rexo);
  const Real lam = G * mMars * mass / (kB * T_exo * r);
  const Real psi = lam * lam / (lam + lc);
  Real frac = (1.0 + P32(lam) - std::sqrt(std::max(1.0 - lam * lam / (lc * lc), 0.0)) * std::exp(-psi) * (1.0 + P32(lam - psi)));
  frac = frac / (1.0 + P32(lc));
  return n0 * frac * std::exp(lam - lc);
}

// d ln nCO2 / dr = -1/H_n (species_density_parameters.cpp:98-99), integrated upward from rmin with the same RK4
Real chamb_diff_1d::CO2_exobase_from_rmin(Real nCO2rmin) const {
  const int nsteps = 400;
  const Real h = (rexo - rmin) / (nsteps - 1);
  auto d = [&](Real r) {
    const Real T = temp.T(r);
    return -(G * mMars * mCO2 / (kB * T * r * r) + temp.Tprime(r) / T);
  };
  Real y = std::log(nCO2rmin);
  for (int i = 0; i < nsteps - 1; i++) {
    const Real r = rmin + i * h;
    y += h / 6.0 * (d(r) + 4 * d(r + 0.5 * h) + d(r + h));   // RK4 of a y-independent right-hand side = Simpson
  }
  return std::exp(y);
}

void chamb_diff_1d::integrate_thermosphere(int nsteps) {
  const Real alpha = species.alpha;
  auto deriv = [&](Real r, const Real (&y)[2], Real (&d)[2]) {
    const Real nCO2 = std::exp(y[0]), nH = std::exp(y[1]);
    const Real T = (r <= rexo) ? temp.T(r) : T_exo;
    const Real Tp = temp.Tprime(r);
    const Real D = std::pow(T, species.s) * species.DH0 / nCO2;
    const Real K = 1.2e12 * std::sqrt(T_exo / nCO2);
    const Real Hn_inv = G * mMars * mCO2 / (kB * T * r * r) + Tp / T;
    const Real HH_inv = G * mMars * species.mass / (kB * T * r * r) + (1 + alpha) * Tp / T;
    d[0] = -Hn_inv;
    d[1] = -(escape_flux * (rexo / r) * (rexo / r) / nH + D * HH_inv + K * Hn_inv) / (D + K);
  };
  std::vector<Real> rs(nsteps);
  const Real step = (rmin - rexo) / (nsteps - 1);
  for (int i = 0; i < nsteps; i++) rs[i] = rexo + i * step;
  rs[nsteps - 1] = rmin;
  const Real h = rs[1] - rs[0];
  Real y[2] = {std::log(nCO2_exo), std::log(nH_exo)};
  std::vector<Real> oc(nsteps), oh(nsteps);
  oc[0] = y[0]; oh[0] = y[1];
  for (int i = 0; i < nsteps - 1; i++) {
    const Real r = rs[i];
    Real k1[2], k2[2], k3[2], k4[2], t[2];
    deriv(r, y, k1);
    t[0] = y[0] + 0.5 * h * k1[0]; t[1] = y[1] + 0.5 * h * k1[1];
    deriv(r + 0.5 * h, t, k2);
    t[0] = y[0] + 0.5 * h * k2[0]; t[1] = y[1] + 0.5 * h * k2[1];
    deriv(r + 0.5 * h, t, k3);
    t[0] = y[0] + h * k3[0]; t[1] = y[1] + h * k3[1];
    deriv(r + h, t, k4);
    y[0] = y[0] + h / 6.0 * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0]);
    y[1] = y[1] + h / 6.0 * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1]);
    oc[i + 1] = y[0]; oh[i + 1] = y[1];
  }
  thermo_r.assign(rs.rbegin(), rs.rend());
  thermo_lnCO2.assign(oc.rbegin(), oc.rend());
  thermo_lnH.assign(oh.rbegin(), oh.rend());
}

Real chamb_diff_1d::n_species(Real r) const {
  if (r >= rexo) return n_exo(r, nH_exo, species.mass);
  if (r >= rmindiffusion) return std::exp(interp(r, thermo_r, thermo_lnH));
  return n_species_rmindiffusion / nCO2_rmindiffusion * n_absorber(r);   // well mixed below rmindiffusion (:213-216)
}

Real chamb_diff_1d::n_absorber(Real r) const {
  if (r >= rexo) {   // CO2 above the exobase: isothermal barometric fall-off, zero above rexo + 500 km
    if (r > rexo + 500e5) return 0.0;
    const Real lam = G * mMars * mCO2 / (kB * T_exo);
    return nCO2_exo * std::exp(lam * (1.0 / std::max(r, rexo) - 1.0 / rexo));
  }
  return std::exp(interp(r, thermo_r, thermo_lnCO2));
}

void chamb_diff_1d::save(const std::string &fname) const {
  std::ofstream file(fname.c_str());
  if (!file.is_open()) return;
  file << "chaffin atmosphere for:\n"
       << "         rexo = " << rexo << " cm,\n"
       << "         Texo = " << T_exo << " K,\n"
       << "n_species_exo = " << nH_exo << " cm-3,\n"
       << "      nCO2exo = " << nCO2_exo << " cm-3,\n\n";
  file << "Thermosphere is defined by solution to Krasnopolsky (2002) differential equation:\n";
  file << "Thermosphere interpolation is log-linear:\n";
  auto dump = [&](const char *pre, const std::vector<Real> &v) {
    file << pre;
    for (size_t i = 0; i < v.size(); i++) file << (i ? " " : "") << v[i];
    file << "\n";
  };
  dump("r [cm] = ", thermo_r);
  dump("log(n_species) [cm-3] = ", thermo_lnH);
  dump("log(nCO2) [cm-3] = ", thermo_lnCO2);
  file << "\nExosphere is spherically symmetric Chamberlain (nCO2 assumed zero):\n";
  file << "Exosphere interpolation is log-log:\n";
  std::vector<Real> lr(100), ln(100);
  for (int i = 0; i < 100; i++) {
    lr[i] = std::log(rexo) + i * (std::log(rmax) - std::log(rexo)) / 99.0;
    ln[i] = std::log(n_species(std::exp(lr[i])));
  }
  dump("logr [cm] = ", lr);
  dump("log(n_species) [cm-3] = ", ln);
}

// ---------------------------------------------------------------- chamb_diff_1d_asymmetric
void chamb_diff_1d_asymmetric::set_asymmetry(Real a) {
  asymmetry = a;
  n0 = 2.0 / (a + 1);
  nslope = 2.0 / pi * (a - 1) / (a + 1);
}

Real chamb_diff_1d_asymmetric::theta_average_factor(Real t0, Real t1) const {   // chamb_diff_1d_asymmetric.cpp:37-51
  const Real tmin = t0 < 0 ? 0 : t0;
  const Real tmax = t1 > pi ? 0 : t1;      // as the reference writes it
  const Real c0 = std::cos(tmin), c1 = std::cos(tmax), s0 = std::sin(tmin), s1 = std::sin(tmax);
  return n0 + nslope * ((s1 - s0 + tmin * c0 - tmax * c1) / (c0 - c1));
}

void chamb_diff_1d_asymmetric::voxel_values(Real r0, Real r1, Real t0, Real t1, Real pt_r, Real pt_t, Real (&out)[6]) const {
  atmosphere_model::voxel_values(r0, r1, t0, t1, pt_r, pt_t, out);
  out[0] *= theta_average_factor(t0, t1);
  out[1] *= nslope * pt_t + n0;
}

// ---------------------------------------------------------------- chamb_diff_temp_asymmetric
chamb_diff_temp_asymmetric::chamb_diff_temp_asymmetric(const species_density_parameters &sp, Real navgg, Real T00, Real T11,
                                                       Real nCO2rminn, Real rexoo, Real rminn, Real rmaxx, Real rmindiffusionn,
                                                       Real T_tropo, Real r_tropo, Real shape_parameter, Real Tpowerr)
    : navg(navgg), T0(T00), T1(T11), Tpower(Tpowerr) {
  rmin = rminn; rexo = rexoo; rmax = rmaxx;
  // normalisation: the sphere average of A T(sza)^-Tpower is navg (trapezoid in sza, weight sin; :62-79)
  const int n_int = 100;
  const Real dth = pi / n_int;
  Real num = 0, den = 0;
  for (int i = 0; i <= n_int; i++) {
    const Real th = i * dth, wgt = (i == 0 || i == n_int) ? 0.5 : 1.0;
    num += wgt * std::pow(T_sza(th), -Tpower) * std::sin(th);
    den += wgt * std::sin(th);
  }
  A = navg * den / num;
  for (int i = 0; i < n_sza; i++) {
    const Real sza = i * pi / (n_sza - 1);
    krasnopolsky_temperature tk(T_sza(sza), T_tropo, r_tropo, shape_parameter, false);
    atm_sza.emplace_back(new chamb_diff_1d(rmin, rexo, rmax, rmindiffusionn, n_species_sza(sza), nCO2rminn, tk, sp,
                                           chamb_diff_1d::method_rmax_nCO2rmin));
  }
}

Real chamb_diff_temp_asymmetric::at(int which, Real r, Real t) const {
  const Real d_sza = pi / (n_sza - 1);
  t = std::min(std::max(t, 0.0), pi);
  int i = (int) (t / d_sza);
  if (i > n_sza - 1) i = n_sza - 1;
  const Real wt = 1.0 - (t - i * d_sza) / d_sza;
  auto f = [&](const chamb_diff_1d &a) { return which == 0 ? a.n_species(r) : which == 1 ? a.Temp(r) : a.n_absorber(r); };
  if (i == n_sza - 1) return f(*atm_sza[i]);
  return wt * f(*atm_sza[i]) + (1.0 - wt) * f(*atm_sza[i + 1]);
}

void chamb_diff_temp_asymmetric::voxel_values(Real r0, Real r1, Real t0, Real t1, Real pt_r, Real pt_t, Real (&out)[6]) const {
  const Real ta = std::max(t0, 0.0), tb = std::min(t1, (Real) pi);
  // volume average with weight r^2 sin(t) (chamb_diff_temp_asymmetric.cpp:143-157): 8-point Gauss-Legendre in sza
  // (the integrand is piecewise linear in sza on a 40-node grid) x the shell average in r
  static std::vector<double> gx, gw;
  if (gx.empty()) gauss_legendre(8, gx, gw);
  for (int which = 0; which < 3; which++) {
    Real num = 0, den = 0;
    for (size_t k = 0; k < gx.size(); k++) {
      const Real t = 0.5 * (tb + ta) + 0.5 * (tb - ta) * gx[k];
      const Real wgt = gw[k] * std::sin(t);
      num += wgt * shell_average([&](Real r) { return at(which, r, t); }, r0, r1);
      den += wgt;
    }
    out[2 * which] = num / den;
    out[2 * which + 1] = at(which, pt_r, pt_t);
  }
  // order of the table is n, n_pt, T, T_pt, nabs, nabs_pt: the loop above wrote species, temperature, absorber
  if (!temp_dependent_sH) out[2] = out[3] = constant_temp_sH;
}

// ---------------------------------------------------------------- tabular_1d
tabular_1d::tabular_1d(Real rminn, Real rexoo, Real rmaxx, bool compute_exospheree) : compute_exosphere(compute_exospheree) {
  rmin = rminn; rexo = rexoo; rmax = rmaxx;
}

void tabular_1d::load_log_species_density(const std::vector<double> &alt, const std::vector<double> &l) {
  if (alt.size() != l.size() || alt.size() < 2) throw std::invalid_argument("tabular_1d: bad species table");
  const bool ascnd = l[1] > l[0];
  for (size_t i = 1; i < l.size(); i++)
    if (ascnd ? !(l[i] > l[i - 1]) : !(l[i] < l[i - 1]))
      throw std::invalid_argument("log_n_species must be monotonic and invertible");
  alt_n = alt; log_n = l;
  check_init();
}
void tabular_1d::load_log_absorber_density(const std::vector<double> &alt, const std::vector<double> &l) {
  if (alt.size() != l.size() || alt.size() < 2) throw std::invalid_argument("tabular_1d: bad absorber table");
  alt_a = alt; log_a = l;
  check_init();
}
void tabular_1d::load_temperature(const std::vector<double> &alt, const std::vector<double> &t) {
  if (alt.size() != t.size() || alt.size() < 2) throw std::invalid_argument("tabular_1d: bad temperature table");
  alt_T = alt; tab_T = t;
  check_init();
}
void tabular_1d::check_init() {
  if (compute_exosphere && !alt_n.empty() && !alt_T.empty()) {
    exo_n0 = std::exp(interp((rexo - rMars) / 1e5, alt_n, log_n));
    exo_T = interp((rexo - rMars) / 1e5, alt_T, tab_T);
    exo_lambdac = G * mMars * m_species / (kB * exo_T * rexo);
  }
}
Real tabular_1d::Temp(Real r) const {
  if (alt_T.empty()) throw std::runtime_error("tabular_1d: Temp must be initialized!");
  if (compute_exosphere && r > rexo) return exo_T;
  return interp((r - rMars) / 1e5, alt_T, tab_T);
}
Real tabular_1d::n_species(Real r) const {
  if (alt_n.empty()) throw std::runtime_error("tabular_1d: n_species must be initialized!");
  if (compute_exosphere && r > rexo) {
    const Real lam = G * mMars * m_species / (kB * exo_T * r);
    const Real psi = lam * lam / (lam + exo_lambdac);
    Real frac = (1.0 + P32(lam) - std::sqrt(std::max(1.0 - lam * lam / (exo_lambdac * exo_lambdac), 0.0)) * std::exp(-psi) * (1.0 + P32(lam - psi)));
    frac = frac / (1.0 + P32(exo_lambdac));
    return exo_n0 * frac * std::exp(lam - exo_lambdac);
  }
  return std::exp(interp((r - rMars) / 1e5, alt_n, log_n));
}
Real tabular_1d::n_absorber(Real r) const {
  if (alt_a.empty()) throw std::runtime_error("tabular_1d: n_absorber must be initialized!");
  if (compute_exosphere && r > rexo) return 0.0;
  return std::exp(interp((r - rMars) / 1e5, alt_a, log_a));
}

// ---------------------------------------------------------------- Temp_converter
double Temp_converter::lc_from_T_exact(double T) const { return G * mMars * m_species / (kB * T * rexo); }
double Temp_converter::eff_from_T_exact(double T) const {
  const double lc = lc_from_T_exact(T);
  return 0.5 * std::sqrt(2.0 * kB * T / (m_species * pi)) * (1.0 + lc) * std::exp(-lc);
}
double Temp_converter::T_from_lc(double lc) const { return G * mMars * m_species / (kB * lc * rexo); }
double Temp_converter::T_from_eff(double eff) const {   // eff is increasing in T on [100, 1200] K
  double lo = 100.0, hi = 1200.0;
  if (eff <= eff_from_T_exact(lo)) return lo;
  if (eff >= eff_from_T_exact(hi)) return hi;
  for (int i = 0; i < 100; i++) {
    const double mid = 0.5 * (lo + hi);
    if (eff_from_T_exact(mid) < eff) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

} // namespace b200rt_host
