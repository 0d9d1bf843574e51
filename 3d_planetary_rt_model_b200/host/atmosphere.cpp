// atmosphere.cpp -- see atmosphere.hpp
#include "atmosphere.hpp"
#include <algorithm>
#include <cmath>

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

chamb_diff_1d::chamb_diff_1d(Real nHexo, Real nCO2exo, Real Texo, Real rmaxx)
    : nH_exo(nHexo), T_exo(Texo), nCO2_exo(nCO2exo) {
  lambdac = G * mMars * mH / (kB * T_exo * rexo);
  const Real veff = 0.5 * std::sqrt(2.0 * kB * T_exo / (mH * pi)) * (1.0 + lambdac) * std::exp(-lambdac);
  escape_flux = nH_exo * veff;
  if (rmaxx > 0) rmax = rmaxx;
  else {
    Real lo = rexo, hi = rexo * 1000.0;
    for (int i = 0; i < 200; i++) {
      const Real mid = std::sqrt(lo * hi);
      if (n_exo(mid) > n_species_min) lo = mid; else hi = mid;
    }
    rmax = 0.5 * (lo + hi);
  }
  integrate_thermosphere();
}

Real chamb_diff_1d::Temp(Real r) const {
  if (r > rexo) return T_exo;
  const Real x = (r - r_tropo) * 1e-5;
  const Real sig = shape * T_exo;
  return (x > 0) ? T_exo - (T_exo - T_tropo) * std::exp(-x * x / sig) : T_tropo;
}

Real chamb_diff_1d::Tprime(Real r) const {
  const Real x = (r - r_tropo) * 1e-5;
  const Real sig = shape * T_exo;
  if (x <= 0) return 0.0;
  const Real T = T_exo - (T_exo - T_tropo) * std::exp(-x * x / sig);
  return (T_exo - T) * (2 * x / sig) * 1e-5;
}

Real chamb_diff_1d::n_exo(Real r) const {
  r = std::max(r, rexo);
  const Real lam = G * mMars * mH / (kB * T_exo * r);
  const Real psi = lam * lam / (lam + lambdac);
  Real frac = (1.0 + P32(lam) - std::sqrt(std::max(1.0 - lam * lam / (lambdac * lambdac), 0.0)) * std::exp(-psi) *
                                    (1.0 + P32(lam - psi)));
  frac = frac / (1.0 + P32(lambdac));
  return nH_exo * frac * std::exp(lam - lambdac);
}

void chamb_diff_1d::integrate_thermosphere(int nsteps) {
  const Real alpha = -0.25;
  auto deriv = [&](Real r, const Real (&y)[2], Real (&d)[2]) {
    const Real nCO2 = std::exp(y[0]), nH = std::exp(y[1]);
    const Real T = (r <= rexo) ? Temp(r) : T_exo;
    const Real Tp = Tprime(r);
    const Real D = std::pow(T, 0.6) * 8.4e17 / nCO2;
    const Real K = 1.2e12 * std::sqrt(T_exo / nCO2);
    const Real Hn_inv = G * mMars * mCO2 / (kB * T * r * r) + Tp / T;
    const Real HH_inv = G * mMars * mH / (kB * T * r * r) + (1 + alpha) * Tp / T;
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
  return (r >= rexo) ? n_exo(r) : std::exp(interp(r, thermo_r, thermo_lnH));
}

Real chamb_diff_1d::n_absorber(Real r) const {
  if (r >= rexo) {   // CO2 above the exobase: isothermal barometric fall-off, zero above rexo + 500 km
    if (r > rexo + 500e5) return 0.0;
    const Real lam = G * mMars * mCO2 / (kB * T_exo);
    return nCO2_exo * std::exp(lam * (1.0 / std::max(r, rexo) - 1.0 / rexo));
  }
  return std::exp(interp(r, thermo_r, thermo_lnCO2));
}

Real chamb_diff_1d::r_from_n_species(Real n) const {
  Real lo = rmin, hi = rmax;
  for (int i = 0; i < 200; i++) {
    const Real mid = 0.5 * (lo + hi);
    if (n_species(mid) > n) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

std::vector<Real> chamb_diff_1d::radial_boundaries(int n_rb, int rmethod) const {
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
Real chamb_diff_1d::shell_average(F f, Real r0, Real r1) const {
  static std::vector<double> gx, gw;
  if (gx.empty()) gauss_legendre(48, gx, gw);
  const Real l0 = std::log(r0), l1 = std::log(r1);
  Real num = 0, den = 0;
  for (size_t k = 0; k < gx.size(); k++) {
    const Real r = std::exp(0.5 * (l1 + l0) + 0.5 * (l1 - l0) * gx[k]);
    const Real jac = r * r * r;   // r^2 dr = r^3 dln r
    num += gw[k] * f(r) * jac;
    den += gw[k] * jac;
  }
  return num / den;
}

void chamb_diff_1d::voxel_tables(const std::vector<Real> &rb, int n_sb, std::vector<Real> (&out)[6]) const {
  const int n_r = (int) rb.size() - 1, n_s = n_sb - 1;
  for (auto &v : out) v.assign((size_t) n_r * n_s, 0.0);
  for (int i = 0; i < n_r; i++) {
    const Real pt = std::sqrt(rb[i] * rb[i + 1]);
    const Real vals[6] = {shell_average([&](Real r) { return n_species(r); }, rb[i], rb[i + 1]), n_species(pt),
                          shell_average([&](Real r) { return Temp(r); }, rb[i], rb[i + 1]), Temp(pt),
                          shell_average([&](Real r) { return n_absorber(r); }, rb[i], rb[i + 1]), n_absorber(pt)};
    for (int j = 0; j < n_s; j++)
      for (int q = 0; q < 6; q++) out[q][(size_t) i * n_s + j] = vals[q];
  }
}

} // namespace b200rt_host
