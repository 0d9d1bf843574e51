// observation_fit.cpp -- see observation_fit.hpp
#include "observation_fit.hpp"
#include <atomic>
#include <chrono>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <thread>
#include "../../include/b200rt.h"

using namespace b200rt_host;
using std::string;
using std::vector;

struct observation_fit::set_inputs {
  int n_rb = 0, n_sb = 0, n_rays = 0, n_vox = 0;
  vector<double> rb, sb, pts_r, pts_s, ray_t, ray_p, ray_w;
  vector<double> vox[6];                 // n_avg, n_pt, T_avg, T_pt, nabs_avg, nabs_pt
  vector<double> tabs[n_hydrogen_emissions][8];
  double branching[n_hydrogen_emissions], T_ref[n_hydrogen_emissions], sigma_ref[n_hydrogen_emissions];
  double abs_sigma[n_hydrogen_emissions];
  // solution, filled after the solve (what save_S writes)
  vector<double> S[n_hydrogen_emissions], S0[n_hydrogen_emissions], tau_sp[n_hydrogen_emissions], tau_abs[n_hydrogen_emissions];
};

void observation_fit::check(int rc, b200rt_ctx *c) const {
  if (rc != B200RT_OK)
    throw std::runtime_error(string("b200rt status ") + std::to_string(rc) + ": " + (c ? b200rt_last_error(c) : "no context"));
}

observation_fit::observation_fit(const string iph_sfn_fnamee, int devicee)
    : device(devicee), iph_sfn_fname(iph_sfn_fnamee) {
  CO2_exobase_density = default_CO2_exobase_density;
  g_factor[0] = lyman_alpha_typical_g_factor;
  g_factor[1] = lyman_beta_typical_g_factor;
  if (b200rt_create(device, B200RT_F64, &ctx) != B200RT_OK)
    throw std::runtime_error("observation_fit: no usable CUDA device (this library has no CPU path)");
  last = new set_inputs;
}

observation_fit::~observation_fit() {
  b200rt_destroy(ctx);
  delete last;
}

void observation_fit::add_observation(const vector<vector<Real>> &MSO_locations, const vector<vector<Real>> &MSO_directions) {
  if (MSO_locations.size() != MSO_directions.size())
    throw std::invalid_argument("location and look direction must have the same length.");
  const int n = (int) MSO_locations.size();
  vector<double> loc(3 * (size_t) n), dir(3 * (size_t) n);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      loc[3 * (size_t) i + k] = MSO_locations[i].at(k);
      dir[3 * (size_t) i + k] = MSO_directions[i].at(k);
    }
  for (auto &a : los) a.assign(n, 0.0);
  int rc = b200rt_los_from_MSO(B200RT_F64, n, loc.data(), dir.data(), los[0].data(), los[1].data(), los[2].data(),
                               los[3].data(), los[4].data(), los[5].data(), los[6].data(), los[7].data(), los[8].data());
  if (rc != B200RT_OK) throw std::runtime_error("b200rt_los_from_MSO failed");
  brightness_done = false;
}

void observation_fit::simulate_iph(const bool sim_iphh) { sim_iph = sim_iphh; }

void observation_fit::add_observation_ra_dec(const vector<Real> &mars_ecliptic_coords, const vector<Real> &RAA,
                                             const vector<Real> &Decc) {
  simulate_iph(true);
  if ((int) RAA.size() != n_obs() || RAA.size() != Decc.size())
    throw std::invalid_argument("IPH coordinates must have the same dimensions as locations and directions");
  if (mars_ecliptic_coords.size() != 3) throw std::invalid_argument("mars ecliptic coords must be a 3D position.");
  mars_ecliptic_pos = mars_ecliptic_coords;
  ra = RAA;
  dec = Decc;
  iph_unextincted.assign(n_obs(), vector<Real>(n_hydrogen_emissions, 0.0));
  iph_observed.assign(n_obs(), vector<Real>(n_hydrogen_emissions, 0.0));
  get_unextincted_iph();
}

void observation_fit::get_unextincted_iph() {
  if (!iph_table_loaded) {
    check(b200rt_iph_load_table(ctx, iph_sfn_fname.c_str()), ctx);
    iph_table_loaded = true;
  }
  vector<double> kR(n_obs());
  check(b200rt_iph_model(ctx, g_factor[0], mars_ecliptic_pos.data(), n_obs(), ra.data(), dec.data(), kR.data()), ctx);
  for (int i = 0; i < n_obs(); i++) {
    iph_unextincted[i][0] = kR[i];
    iph_unextincted[i][1] = g_factor[1] / g_factor[0] * kR[i];   // observation_fit.cpp:114-117
  }
  brightness_done = false;
}

void observation_fit::set_g_factor(vector<Real> &g) {
  for (int e = 0; e < n_hydrogen_emissions; e++) {
    g_factor[e] = g.at(e);
    b200rt_set_g_factor(ctx, e, g_factor[e]);
  }
  std::cout << "Ly alpha solar brightness = " << g[0] / lyman_alpha_cross_section_total << std::endl;
  std::cout << "Ly beta solar brightness = " << g[1] / lyman_beta_cross_section_total << std::endl;
  brightness_done = false;
}

// ---- one parameter set: atmosphere -> grid -> the tables of singlet_CFR::define (singlet_CFR.hpp:419-492)
void observation_fit::build_inputs(const Real &nHexo, const Real &Texo, set_inputs &in) const {
  chamb_diff_1d atm(nHexo, CO2_exobase_density, Texo);
  static_cast<H_cross_sections &>(atm) = H_cross_section_options;       // atm.copy_H_options
  in.n_rb = n_radial_boundaries; in.n_sb = n_sza_boundaries; in.n_rays = n_rays_theta * n_rays_phi;
  in.n_vox = n_voxels;
  in.rb = atm.radial_boundaries(n_radial_boundaries, 0);                  // rmethod_altitude, observation_fit.cpp:34
  in.sb.assign(in.n_sb, 0); in.pts_r.assign(in.n_rb - 1, 0); in.pts_s.assign(in.n_sb - 1, 0);
  in.ray_t.assign(in.n_rays, 0); in.ray_p.assign(in.n_rays, 0); in.ray_w.assign(in.n_rays, 0);
  int rc = b200rt_make_grid_sph(B200RT_F64, in.n_rb, in.n_sb, n_rays_theta, n_rays_phi, in.rb.data(), szamethod,
                                1 /* raymethod_theta_uniform */, in.sb.data(), in.pts_r.data(), in.pts_s.data(),
                                in.ray_t.data(), in.ray_p.data(), in.ray_w.data());
  if (rc != B200RT_OK) throw std::runtime_error("b200rt_make_grid_sph failed");
  atm.voxel_tables(in.rb, in.n_sb, in.vox);
  const double br[2] = {1.0, lyman_beta_branching_ratio};
  const double sref[2] = {atm.sH_lya(Texo), atm.sH_lyb(Texo)};
  for (int e = 0; e < n_hydrogen_emissions; e++) {
    in.branching[e] = br[e]; in.T_ref[e] = Texo; in.sigma_ref[e] = sref[e];
    for (auto &t : in.tabs[e]) t.assign(in.n_vox, 0.0);
    for (int v = 0; v < in.n_vox; v++) {
      const double n_avg = in.vox[0][v], n_pt = in.vox[1][v], T_avg = in.vox[2][v], T_pt = in.vox[3][v];
      const double a_avg = in.vox[4][v], a_pt = in.vox[5][v];
      const double Tr = Texo / T_avg, Tr_pt = Texo / T_pt;
      const double sa = (e == 0) ? atm.sCO2_lya(T_avg) : atm.sCO2_lyb(T_avg);
      const double sa_pt = (e == 0) ? atm.sCO2_lya(T_pt) : atm.sCO2_lyb(T_pt);
      in.abs_sigma[e] = sa;
      in.tabs[e][0][v] = Tr;                             // species_T_ratio
      in.tabs[e][1][v] = n_avg;                          // species_density
      in.tabs[e][2][v] = n_avg * sref[e] * std::sqrt(Tr);         // dtau_species
      in.tabs[e][3][v] = a_avg * sa;                     // dtau_absorber
      in.tabs[e][4][v] = Tr_pt;
      in.tabs[e][5][v] = n_pt;
      in.tabs[e][6][v] = n_pt * sref[e] * std::sqrt(Tr_pt);
      in.tabs[e][7][v] = a_pt * sa_pt;
    }
  }
}

void observation_fit::load_inputs(b200rt_ctx *c, const set_inputs &in) const {
  check(b200rt_set_grid_sph(c, in.n_rb, in.n_sb, in.n_rays, in.rb.data(), in.sb.data(), in.pts_r.data(), in.pts_s.data(),
                            in.ray_t.data(), in.ray_p.data(), in.ray_w.data()), c);
  for (int e = 0; e < n_hydrogen_emissions; e++)
    check(b200rt_set_singlet(c, e, n_hydrogen_emissions, in.branching[e], in.T_ref[e], in.sigma_ref[e], g_factor[e],
                             in.tabs[e][0].data(), in.tabs[e][1].data(), in.tabs[e][2].data(), in.tabs[e][3].data(),
                             in.tabs[e][4].data(), in.tabs[e][5].data(), in.tabs[e][6].data(), in.tabs[e][7].data()), c);
}

void observation_fit::generate_source_function(const Real &nHexo, const Real &Texo, const string atmosphere_fname,
                                               const string sourcefn_fname, const bool plane_parallel, const bool deuterium) {
  if (plane_parallel) throw std::runtime_error("observation_fit: the plane-parallel grid is not built in this facade");
  if (deuterium) throw std::runtime_error("observation_fit: the deuterium model is not built in this facade");
  if (atmosphere_fname != "") throw std::runtime_error("observation_fit: atm.save is not built in this facade");
  build_inputs(nHexo, Texo, *last);
  load_inputs(ctx, *last);
  check(b200rt_generate_S(ctx), ctx);            // RT_obj.generate_S_gpu(), observation_fit.hpp:286-290
  rb = last->rb;
  for (int e = 0; e < n_hydrogen_emissions; e++) {
    for (auto *v : {&last->S[e], &last->S0[e], &last->tau_sp[e], &last->tau_abs[e]}) v->assign(n_voxels, 0.0);
    check(b200rt_get_solution(ctx, e, last->S[e].data(), last->S0[e].data(), last->tau_sp[e].data(), last->tau_abs[e].data()), ctx);
  }
  have_S = true;
  brightness_done = false;
  if (sourcefn_fname != "") save_S(sourcefn_fname, *last);
}

vector<observation_fit::Real> observation_fit::source_function(int e) {
  if (!have_S) throw std::runtime_error("observation_fit: no source function yet");
  return last->S[e];
}

// RT_grid::brightness_gpu(obs) + the packing of observation_fit.cpp:491-559
void observation_fit::run_brightness(b200rt_ctx *c, bool upload, vector<vector<Real>> (&q)[4]) const {
  const int n = n_obs();
  vector<double> flat[4];
  for (auto &f : flat) f.assign((size_t) n_hydrogen_emissions * n, 0.0);
  if (upload)
    check(b200rt_los_upload(c, n, los[0].data(), los[1].data(), los[2].data(), los[3].data(), los[4].data(),
                            los[5].data(), los[6].data(), los[7].data(), los[8].data()), c);
  check(b200rt_brightness_resident(c, 10), c);
  check(b200rt_los_download(c, flat[0].data(), flat[1].data(), flat[2].data(), flat[3].data()), c);
  for (int k = 0; k < 4; k++) {
    q[k].assign(n_hydrogen_emissions, vector<Real>());
    for (int e = 0; e < n_hydrogen_emissions; e++) q[k][e].assign(flat[k].begin() + (size_t) e * n, flat[k].begin() + (size_t) (e + 1) * n);
  }
}

vector<vector<observation_fit::Real>> observation_fit::brightness() {
  if (!have_S) throw std::runtime_error("observation_fit: generate a source function before asking for brightness");
  if (n_obs() == 0) throw std::runtime_error("there must be at least one observation to simulate");
  if (!brightness_done) {
    run_brightness(ctx, true, out_q);
    if (sim_iph)                                  // observation::update_iph_extinction, observation.hpp:144-154
      for (int i = 0; i < n_obs(); i++)
        for (int e = 0; e < n_hydrogen_emissions; e++) {
          const Real ta = out_q[2][e][i];
          iph_observed[i][e] = (ta != -1) ? iph_unextincted[i][e] * std::exp(-ta) : 0.0;
        }
    brightness_done = true;
  }
  vector<vector<Real>> b = out_q[0];
  if (sim_iph)
    for (int e = 0; e < n_hydrogen_emissions; e++)
      for (int i = 0; i < n_obs(); i++) b[e][i] += iph_observed[i][e];
  return b;
}

vector<vector<observation_fit::Real>> observation_fit::species_col_dens() { brightness(); return out_q[3]; }
vector<vector<observation_fit::Real>> observation_fit::tau_species_final() { brightness(); return out_q[1]; }
vector<vector<observation_fit::Real>> observation_fit::tau_absorber_final() { brightness(); return out_q[2]; }
vector<vector<observation_fit::Real>> observation_fit::iph_brightness_observed() {
  brightness();
  vector<vector<Real>> r(n_hydrogen_emissions, vector<Real>(n_obs(), 0.0));   // [i_emission][i_obs], observation_fit.cpp:561-575
  if (sim_iph)
    for (int e = 0; e < n_hydrogen_emissions; e++)
      for (int i = 0; i < n_obs(); i++) r[e][i] = iph_observed[i][e];
  return r;
}
vector<vector<observation_fit::Real>> observation_fit::iph_brightness_unextincted() {
  vector<vector<Real>> r(n_hydrogen_emissions, vector<Real>(n_obs(), 0.0));
  if (sim_iph)
    for (int e = 0; e < n_hydrogen_emissions; e++)
      for (int i = 0; i < n_obs(); i++) r[e][i] = iph_unextincted[i][e];
  return r;
}

// ---- options (observation_fit.cpp:413-487)
void observation_fit::set_use_CO2_absorption(const bool use) { H_cross_section_options.no_CO2_absorption = !use; }
void observation_fit::set_use_temp_dependent_sH(const bool use, const Real constant_temp_sH) {
  H_cross_section_options.temp_dependent_sH = use;
  H_cross_section_options.constant_temp_sH = constant_temp_sH;
  if (!use && constant_temp_sH == -1) throw std::invalid_argument("set_use_temp_dependent_sH: a constant temperature is needed");
}
void observation_fit::set_sza_method_uniform() { szamethod = 0; }
void observation_fit::set_sza_method_uniform_cos() { szamethod = 1; }
void observation_fit::reset_H_lya_xsec_coef(const Real x) { H_cross_section_options.H_lya_xsec_coef = x; }
void observation_fit::reset_H_lyb_xsec_coef(const Real x) { H_cross_section_options.H_lyb_xsec_coef = x; }
void observation_fit::reset_CO2_lya_xsec(const Real x) { H_cross_section_options.CO2_lya_xsec = x; }
void observation_fit::reset_CO2_lyb_xsec(const Real x) { H_cross_section_options.CO2_lyb_xsec = x; }
observation_fit::Real observation_fit::get_CO2_exobase_density() { return CO2_exobase_density; }
void observation_fit::reset_CO2_exobase_density() { CO2_exobase_density = default_CO2_exobase_density; }
void observation_fit::set_CO2_exobase_density(const double nCO2) { CO2_exobase_density = nCO2; }

// ---- ASCII writers (grid_spherical_azimuthally_symmetric.hpp:630-665, singlet_CFR.hpp:519-543,
// emission_voxels.hpp:235-238).  Numbers are printed the way Eigen's default IOFormat prints a
// row vector: stream precision 6, every coefficient right-aligned to the widest one, one space between.
namespace {
string row(const vector<double> &v) {
  vector<string> s(v.size());
  size_t w = 0;
  for (size_t i = 0; i < v.size(); i++) {
    std::ostringstream o;
    o << v[i];
    s[i] = o.str();
    w = std::max(w, s[i].size());
  }
  std::ostringstream o;
  for (size_t i = 0; i < v.size(); i++) {
    if (i) o << " ";
    o << std::setw((int) w) << s[i];
  }
  return o.str();
}
}

void observation_fit::save_S(const string &fname, const set_inputs &in) {
  std::ofstream file(fname.c_str());
  if (!file.is_open()) return;
  file << "radial boundaries [cm]: " << row(in.rb) << "\n\n";
  file << "pts radii [cm]: " << row(in.pts_r) << "\n\n";
  file << "sza boundaries [rad]: " << row(in.sb) << "\n\n";
  file << "pts sza [rad]: " << row(in.pts_s) << "\n\n";
  const char *names[2] = {"H Lyman alpha", "H Lyman beta"};
  const int ns = in.n_sb - 1, nr = in.n_rb - 1;
  auto slice = [&](const vector<double> &q, int j) {      // sza_slice: every voxel of SZA column j
    vector<double> r(nr);
    for (int i = 0; i < nr; i++) r[i] = q[(size_t) i * ns + j];
    return r;
  };
  for (int e = 0; e < n_hydrogen_emissions; e++) {
    file << "For " << names[e] << "\n";
    vector<double> sig(in.n_vox), asig(in.n_vox, in.abs_sigma[e]);
    for (int v = 0; v < in.n_vox; v++) sig[v] = in.sigma_ref[e] * std::sqrt(in.tabs[e][0][v]);
    for (int j = 0; j < ns; j++) {
      file << "  For SZA = " << in.pts_s[j] << ": \n";
      file << "    Species density [cm-3]: " << row(slice(in.tabs[e][1], j)) << "\n"
           << "    Species single scattering tau: " << row(slice(in.tau_sp[e], j)) << "\n"
           << "    Species cross section [cm2]: " << row(slice(sig, j)) << "\n"
           << "    Absorber density [cm-3]: " << row(slice(in.vox[4], j)) << "\n"
           << "    Absorber single scattering tau: " << row(slice(in.tau_abs[e], j)) << "\n"
           << "    Absorber cross section [cm2]: " << row(slice(asig, j)) << "\n"
           << "    Species single scattering source function S0: " << row(slice(in.S0[e], j)) << "\n"
           << "    Source function: " << row(slice(in.S[e], j)) << "\n\n";
    }
  }
}

void observation_fit::save_influence_matrix(const string fname) {
  if (!have_S) throw std::runtime_error("observation_fit: no influence matrix yet");
  std::ofstream file(fname.c_str());
  if (!file.is_open()) return;
  const char *names[2] = {"H Lyman alpha", "H Lyman beta"};
  vector<double> K((size_t) n_voxels * n_voxels);
  for (int e = 0; e < n_hydrogen_emissions; e++) {
    check(b200rt_get_influence(ctx, e, B200RT_ROW_MAJOR, K.data()), ctx);   // fetched lazily: K lives on the device
    file << "Here is the influence matrix for " << names[e] << ":\n";
    // Eigen prints a matrix with every coefficient padded to the widest one of the whole matrix
    vector<string> s(K.size());
    size_t w = 0;
    for (size_t i = 0; i < K.size(); i++) {
      std::ostringstream o;
      o << K[i];
      s[i] = o.str();
      w = std::max(w, s[i].size());
    }
    for (int r = 0; r < n_voxels; r++) {
      for (int cidx = 0; cidx < n_voxels; cidx++) {
        if (cidx) file << " ";
        file << std::setw((int) w) << s[(size_t) r * n_voxels + cidx];
      }
      file << "\n";
    }
    file << "\n";
  }
}

// ---- the sweep: every parameter set is independent (own atmosphere, own grid: rmax depends on (nH, T)),
// so sets are handed to worker threads, each owning one context (stream) on one GPU
vector<vector<vector<observation_fit::Real>>> observation_fit::brightness_batch(const vector<Real> &nHexo, const vector<Real> &Texo,
                                                                               int contexts_per_gpu, int n_gpus) {
  if (nHexo.size() != Texo.size()) throw std::invalid_argument("brightness_batch: nHexo and Texo must have the same length");
  if (n_obs() == 0) throw std::runtime_error("there must be at least one observation to simulate");
  const int n_sets = (int) nHexo.size();
  if (n_gpus <= 0) n_gpus = b200rt_device_count();
  n_gpus = std::max(1, std::min(n_gpus, b200rt_device_count()));
  const int n_workers = std::max(1, std::min(n_sets, n_gpus * std::max(1, contexts_per_gpu)));
  vector<vector<vector<Real>>> result(n_sets);
  std::atomic<int> next(0);
  vector<string> errors(n_workers);
  const auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int w) {
    b200rt_ctx *c = nullptr;
    try {
      if (b200rt_create(w % n_gpus, B200RT_F64, &c) != B200RT_OK) throw std::runtime_error("b200rt_create failed");
      bool first = true;
      set_inputs in;
      for (int i = next++; i < n_sets; i = next++) {
        build_inputs(nHexo[i], Texo[i], in);
        load_inputs(c, in);
        check(b200rt_generate_S(c), c);
        vector<vector<Real>> q[4];
        run_brightness(c, first, q);
        first = false;
        if (sim_iph)
          for (int e = 0; e < n_hydrogen_emissions; e++)
            for (int k = 0; k < n_obs(); k++) {
              const Real ta = q[2][e][k];
              q[0][e][k] += (ta != -1) ? iph_unextincted[k][e] * std::exp(-ta) : 0.0;
            }
        result[i] = std::move(q[0]);
      }
    } catch (const std::exception &ex) {
      errors[w] = ex.what();
    }
    b200rt_destroy(c);
  };
  vector<std::thread> pool;
  for (int w = 0; w < n_workers; w++) pool.emplace_back(work, w);
  for (auto &t : pool) t.join();
  batch_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto &e : errors)
    if (!e.empty()) throw std::runtime_error("brightness_batch: " + e);
  return result;
}
