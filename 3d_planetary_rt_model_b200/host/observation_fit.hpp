// observation_fit.hpp -- host facade with the public interface of the reference's observation_fit
// (src/observation_fit.hpp:23-389) for the H Lyman alpha / Lyman beta path, implemented on the C ABI
// of include/b200rt.h.  Plain C++ (g++); nothing here needs nvcc.
//
// What each member maps to (reference src/):
//   add_observation            observation_fit.cpp:62-66 + observation::add_MSO_observation (observation.hpp:46-65)
//   set_g_factor               observation_fit.cpp:74-92
//   simulate_iph, add_observation_ra_dec, get_unextincted_iph        observation_fit.cpp:94-119
//   generate_source_function   observation_fit.cpp:122-169 -> generate_source_function_sph_azi_sym (hpp:241-294)
//   brightness & friends       observation_fit.cpp:491-559, 610-637
//   save_influence_matrix      observation_fit.cpp:647-649
//   set_use_CO2_absorption ... observation_fit.cpp:413-487
// New: brightness_batch() runs a whole (nH, T) sweep -- one atmosphere, grid,
// source function and brightness per parameter set -- over all GPUs with several contexts each,
// instead of the reference's Python loop over generate_source_function + brightness.
//
// Not built in this facade (the calls throw std::runtime_error): plane_parallel, deuterium,
// asymmetric / tabular atmospheres, the O I 102.6 and Lyman multiplet models.
#pragma once
#include <string>
#include <vector>
#include "atmosphere.hpp"

struct b200rt_ctx;

class observation_fit {
public:
  typedef b200rt_host::Real Real;
  static const int n_radial_boundaries = 40;   // observation_fit.hpp:44-47 (standard resolution case)
  static const int n_sza_boundaries = 20;
  static const int n_rays_theta = 7;
  static const int n_rays_phi = 12;
  static const int n_hydrogen_emissions = 2;
  static const int n_voxels = (n_radial_boundaries - 1) * (n_sza_boundaries - 1);

  explicit observation_fit(const std::string iph_sfn_fnamee, int device = 0);
  ~observation_fit();
  observation_fit(const observation_fit &) = delete;
  observation_fit &operator=(const observation_fit &) = delete;

  void add_observation(const std::vector<std::vector<Real>> &MSO_locations,
                       const std::vector<std::vector<Real>> &MSO_directions);
  void get_unextincted_iph();
  void set_g_factor(std::vector<Real> &g);
  void simulate_iph(const bool sim_iphh);
  void add_observation_ra_dec(const std::vector<Real> &mars_ecliptic_coords, const std::vector<Real> &RAA,
                              const std::vector<Real> &Decc);

  void generate_source_function(const Real &nHexo, const Real &Texo, const std::string atmosphere_fname = "",
                                const std::string sourcefn_fname = "", const bool plane_parallel = false,
                                const bool deuterium = false);

  void set_use_CO2_absorption(const bool use_CO2_absorption = true);
  void set_use_temp_dependent_sH(const bool use_temp_dependent_sH = true, const Real constant_temp_sH = -1);
  void set_sza_method_uniform();
  void set_sza_method_uniform_cos();
  void reset_H_lya_xsec_coef(const Real xsec_coef = b200rt_host::lyman_alpha_line_center_cross_section_coef);
  void reset_H_lyb_xsec_coef(const Real xsec_coef = b200rt_host::lyman_beta_line_center_cross_section_coef);
  void reset_CO2_lya_xsec(const Real xsec = b200rt_host::CO2_lyman_alpha_absorption_cross_section);
  void reset_CO2_lyb_xsec(const Real xsec = b200rt_host::CO2_lyman_beta_absorption_cross_section);
  Real get_CO2_exobase_density();
  void reset_CO2_exobase_density();
  void set_CO2_exobase_density(const double nCO2);

  void save_influence_matrix(const std::string fname);

  std::vector<std::vector<Real>> brightness();
  std::vector<std::vector<Real>> species_col_dens();
  std::vector<std::vector<Real>> tau_species_final();
  std::vector<std::vector<Real>> tau_absorber_final();
  std::vector<std::vector<Real>> iph_brightness_observed();
  std::vector<std::vector<Real>> iph_brightness_unextincted();

  // ---- batched sweep: result[i_set][i_emission][i_obs] = brightness (incl. extincted IPH if simulated)
  std::vector<std::vector<std::vector<Real>>> brightness_batch(const std::vector<Real> &nHexo, const std::vector<Real> &Texo,
                                                               int contexts_per_gpu = 4, int n_gpus = -1);
  double last_batch_seconds() const { return batch_seconds; }

  // ---- access to the solution of the last generate_source_function (what save_S writes)
  std::vector<Real> source_function(int i_emission);
  std::vector<Real> radial_boundaries() const { return rb; }
  int n_obs() const { return (int) los[0].size(); }

private:
  struct set_inputs;   // grid + the eight tables of each emission for one parameter set
  void build_inputs(const Real &nHexo, const Real &Texo, set_inputs &in) const;
  void load_inputs(b200rt_ctx *c, const set_inputs &in) const;
  void run_brightness(b200rt_ctx *c, bool upload, std::vector<std::vector<Real>> (&q)[4]) const;
  void check(int rc, b200rt_ctx *c) const;
  void save_S(const std::string &fname, const set_inputs &in);

  int device;
  b200rt_ctx *ctx = nullptr;
  b200rt_host::H_cross_sections H_cross_section_options;
  const Real default_CO2_exobase_density = 2e8;
  Real CO2_exobase_density;
  int szamethod = 1;                   // szamethod_uniform_cos (observation_fit.cpp:37)
  Real g_factor[n_hydrogen_emissions];
  std::string iph_sfn_fname;
  bool sim_iph = false, iph_table_loaded = false;

  // observation (observation.hpp): geometry as the nine atmo_vector fields + the tracker outputs
  std::vector<Real> los[9];
  std::vector<Real> mars_ecliptic_pos, ra, dec;
  std::vector<std::vector<Real>> iph_unextincted, iph_observed;   // [i_obs][i_emission]
  std::vector<std::vector<Real>> out_q[4];                        // [quantity][i_emission][i_obs]
  bool have_S = false, brightness_done = false;
  std::vector<Real> rb;
  set_inputs *last = nullptr;
  double batch_seconds = 0;
};
