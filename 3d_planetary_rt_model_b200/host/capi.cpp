// capi.cpp -- flat C handles over the observation_fit facade, for ctypes-driven tests and for bindings
// that cannot consume C++ (the reference's own binding is Cython on the C++ class, python/py_corona_sim.pyx).
#include <csignal>
#include <cstdlib>
#include <cstring>
#include <execinfo.h>
#include <unistd.h>
#include <string>
#include <vector>
#include "observation_fit.hpp"

namespace {
thread_local std::string g_err;
template <class F> int guard(F f) {
  try { f(); return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
std::vector<std::vector<double>> rows3(int n, const double *a) {
  std::vector<std::vector<double>> r(n, std::vector<double>(3));
  for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) r[i][k] = a[3 * (size_t) i + k];
  return r;
}
void flatten(const std::vector<std::vector<double>> &v, double *out) {
  size_t p = 0;
  for (auto &r : v) { std::memcpy(out + p, r.data(), r.size() * sizeof(double)); p += r.size(); }
}
}

namespace {
// B200RT_HOST_BACKTRACE=1: print a native backtrace on SIGSEGV (development aid)
void segv_handler(int sig) {
  void *frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "libb200rt_host: fatal signal, native backtrace:\n";
  (void) !write(2, msg, sizeof msg - 1);
  backtrace_symbols_fd(frames, n, 2);
  _exit(128 + sig);
}
struct install_handler {
  install_handler() { if (getenv("B200RT_HOST_BACKTRACE")) signal(SIGSEGV, segv_handler); }
} g_install_handler;
}

extern "C" {
const char *obsfit_last_error() { return g_err.c_str(); }
void *obsfit_create(const char *iph_fname, int device) {
  observation_fit *o = nullptr;
  if (guard([&] { o = new observation_fit(iph_fname ? iph_fname : "", device); })) return nullptr;
  return o;
}
void obsfit_destroy(void *h) { delete static_cast<observation_fit *>(h); }
int obsfit_add_observation(void *h, int n, const double *loc, const double *dir) {
  return guard([&] { static_cast<observation_fit *>(h)->add_observation(rows3(n, loc), rows3(n, dir)); });
}
int obsfit_set_g_factor(void *h, double g_lya, double g_lyb) {
  return guard([&] { std::vector<double> g = {g_lya, g_lyb}; static_cast<observation_fit *>(h)->set_g_factor(g); });
}
int obsfit_add_observation_ra_dec(void *h, const double *marspos, int n, const double *ra, const double *dec) {
  return guard([&] {
    static_cast<observation_fit *>(h)->add_observation_ra_dec(std::vector<double>(marspos, marspos + 3),
                                                              std::vector<double>(ra, ra + n), std::vector<double>(dec, dec + n));
  });
}
int obsfit_generate_source_function(void *h, double nH, double T, const char *sourcefn_fname) {
  return guard([&] { static_cast<observation_fit *>(h)->generate_source_function(nH, T, "", sourcefn_fname ? sourcefn_fname : ""); });
}
int obsfit_set_use_CO2_absorption(void *h, int use) { return guard([&] { static_cast<observation_fit *>(h)->set_use_CO2_absorption(use != 0); }); }
int obsfit_set_CO2_exobase_density(void *h, double n) { return guard([&] { static_cast<observation_fit *>(h)->set_CO2_exobase_density(n); }); }
int obsfit_save_influence_matrix(void *h, const char *fname) { return guard([&] { static_cast<observation_fit *>(h)->save_influence_matrix(fname); }); }
// which: 0 brightness, 1 species_col_dens, 2 tau_species_final, 3 tau_absorber_final, 4 iph observed, 5 iph unextincted
int obsfit_get(void *h, int which, double *out) {
  return guard([&] {
    auto *o = static_cast<observation_fit *>(h);
    switch (which) {
      case 0: flatten(o->brightness(), out); break;
      case 1: flatten(o->species_col_dens(), out); break;
      case 2: flatten(o->tau_species_final(), out); break;
      case 3: flatten(o->tau_absorber_final(), out); break;
      case 4: flatten(o->iph_brightness_observed(), out); break;
      default: flatten(o->iph_brightness_unextincted(), out); break;
    }
  });
}
int obsfit_source_function(void *h, int e, double *out) {
  return guard([&] { auto s = static_cast<observation_fit *>(h)->source_function(e); std::memcpy(out, s.data(), s.size() * sizeof(double)); });
}
int obsfit_radial_boundaries(void *h, double *out) {
  return guard([&] { auto s = static_cast<observation_fit *>(h)->radial_boundaries(); std::memcpy(out, s.data(), s.size() * sizeof(double)); });
}
// out[n_sets][2][n_obs]
int obsfit_brightness_batch(void *h, int n_sets, const double *nH, const double *T, int contexts_per_gpu, int n_gpus,
                            double *out, double *seconds) {
  return guard([&] {
    auto *o = static_cast<observation_fit *>(h);
    auto r = o->brightness_batch(std::vector<double>(nH, nH + n_sets), std::vector<double>(T, T + n_sets), contexts_per_gpu, n_gpus);
    size_t p = 0;
    for (auto &s : r) for (auto &e : s) { std::memcpy(out + p, e.data(), e.size() * sizeof(double)); p += e.size(); }
    if (seconds) *seconds = o->last_batch_seconds();
  });
}
// the atmosphere on its own (host only): [6][n_vox] tables and the radial boundaries for (nH, T)
int obsfit_atmosphere_tables(double nH, double nCO2, double T, int n_rb, int n_sb, int rmethod, double *rb_out, double *tables_out) {
  return guard([&] {
    b200rt_host::chamb_diff_1d atm(nH, nCO2, T);
    auto rb = atm.radial_boundaries(n_rb, rmethod);
    std::vector<double> t[6];
    atm.voxel_tables(rb, n_sb, t);
    std::memcpy(rb_out, rb.data(), rb.size() * sizeof(double));
    for (int q = 0; q < 6; q++) std::memcpy(tables_out + (size_t) q * t[q].size(), t[q].data(), t[q].size() * sizeof(double));
  });
}
}
