// atmosphere.hpp -- Boost-free restatement of the reference's atmosphere models (src/atm/*; SURVEY.md appendix G and
// section 8(f) N2 / N4): the host-side INPUT GENERATORS of observation_fit::generate_source_function*.  Each runs once
// per parameter set and feeds per-voxel arrays to the device path.
//
//   chamb_diff_1d               chamb_diff_1d.*, thermosphere_exosphere.*, chamberlain_exosphere.*, temperature.*,
//                               species_density_parameters.*, atmosphere_average_1d.*
//   chamb_diff_1d_asymmetric    chamb_diff_1d_asymmetric.*       (linear-in-SZA density asymmetry)
//   chamb_diff_temp_asymmetric  chamb_diff_temp_asymmetric.*     (exobase temperature linear in SZA, n T^p = const)
//   tabular_1d                  tabular_1d.*, tabular_atmosphere.*
//
// The reference builds these on Boost (gamma_p, odeint, cubic B-splines, root bracketing), which is not available;
// parity of the hot path is evaluated at the voxel-array boundary, so these classes are faithful
// physical restatements with their own numerics:
//   temperature  Krasnopolsky profile                                              temperature.cpp:23-48
//   exosphere    Chamberlain, P(3/2,x) = erf(sqrt x) - 2 sqrt(x/pi) e^-x           chamberlain_exosphere.cpp:22-59
//   thermosphere diffusive species in CO2, classic RK4 from the exobase down       species_density_parameters.cpp:83-156
//   averages     shell averages by Gauss-Legendre quadrature (r^2 weight on the spherical grid, flat weight on the
//                plane-parallel one: atmosphere_average_1d.cpp:141-157)
// chamb_diff_1d with its default arguments is numerically the same algorithm as the Python generator the tests use
// (3d_planetary_rt_model_b200/synth.py), so that the facade and the oracle pipeline can be compared end to end.
#pragma once
#include <memory>
#include <vector>
#include "constants.hpp"

namespace b200rt_host {

// cross sections and their switches (reference src/atm/hydrogen_cross_sections.{hpp,cpp})
struct H_cross_sections {
  double H_lya_xsec_coef = lyman_alpha_line_center_cross_section_coef;
  double H_lyb_xsec_coef = lyman_beta_line_center_cross_section_coef;
  double CO2_lya_xsec = CO2_lyman_alpha_absorption_cross_section;
  double CO2_lyb_xsec = CO2_lyman_beta_absorption_cross_section;
  bool temp_dependent_sH = true;
  double constant_temp_sH = -1;
  bool no_CO2_absorption = false;
  Real sH_lya(const Real &T) const { return H_lya_xsec_coef / std::sqrt(temp_dependent_sH ? T : constant_temp_sH); }
  Real sH_lyb(const Real &T) const { return H_lyb_xsec_coef / std::sqrt(temp_dependent_sH ? T : constant_temp_sH); }
  Real sCO2_lya(const Real &) const { return no_CO2_absorption ? 0.0 : CO2_lya_xsec; }
  Real sCO2_lyb(const Real &) const { return no_CO2_absorption ? 0.0 : CO2_lyb_xsec; }
  void copy_H_options(const H_cross_sections &o) { *this = o; }
};

// temperature.hpp:20-47, temperature.cpp:23-48
struct krasnopolsky_temperature {
  double T_exo, T_tropo, r_tropo, shape_parameter;   // T = T_exo - (T_exo - T_tropo) exp(-x^2 / shape_parameter), x in km
  krasnopolsky_temperature(double T_exoo = 200.0, double T_tropoo = 125.0, double r_tropoo = rMars + 90e5,
                           double shape_parameterr = 11.4, bool shape_parameter_Texo = true);
  double T(double r) const;
  double Tprime(double r) const;
};

// species_density_parameters.hpp: mass, thermal diffusion coefficient, D = T^s DH0 / nCO2
struct species_density_parameters {
  double mass, alpha, DH0, s;
  static species_density_parameters hydrogen() { return {mH, -0.25, 8.4e17, 0.6}; }
  static species_density_parameters deuterium() { return {2.0 * mH, -0.25, 8.4e17, 0.6}; }
  static species_density_parameters oxygen() { return {16.0 * mH, 0.0, 4.4e17, 0.5}; }   // the O source term is 0 there (:244)
};

// what every model offers the grid and the emissions (atmosphere_base.hpp, atmosphere_average_1d.hpp)
class atmosphere_model : public H_cross_sections {
public:
  Real rmin = 0, rexo = 0, rmax = 0;
  bool spherical = true;          // averaging weight: r^2 (spherical grid) or flat (plane parallel)
  virtual ~atmosphere_model() {}

  virtual Real Temp(Real r) const = 0;
  virtual Real n_species(Real r) const = 0;
  virtual Real n_absorber(Real r) const = 0;
  virtual Real r_from_n_species(Real n) const;     // default: bisection on [rmin, rmax]

  // radial boundaries: rmethod 0 = altitude (coordinate_generation.hpp:57-87), 1 = log n_species
  // (grid_spherical_azimuthally_symmetric.hpp:178-187, grid_plane_parallel.hpp:141-160)
  std::vector<Real> radial_boundaries(int n_rb, int rmethod) const;

  // per-voxel inputs of singlet_CFR::define (singlet_CFR.hpp:419-492): out = n_avg, n_pt, T_avg, T_pt, nabs_avg, nabs_pt
  // for the voxel [r0, r1] x [t0, t1] with point (pt_r, pt_t).  Default: the 1-D profile (no SZA dependence);
  // T is the constant temperature when temp_dependent_sH is off (chamb_diff_1d.cpp Temp_voxel_avg).
  virtual void voxel_values(Real r0, Real r1, Real t0, Real t1, Real pt_r, Real pt_t, Real (&out)[6]) const;
  virtual bool sza_dependent() const { return false; }   // false: voxel_tables evaluates one SZA column and copies it

  // [6][n_vox] tables, voxel id = ir*(n_sb-1)+isza.  sb / pts_s = SZA boundaries / points of the grid
  // (n_sb = 2, sb = {0, pi} on the plane-parallel grid).
  void voxel_tables(const std::vector<Real> &rb, const std::vector<Real> &pts_r, const std::vector<Real> &sb,
                    const std::vector<Real> &pts_s, std::vector<Real> (&out)[6]) const;
  // the same for a 1-D model with points at sqrt(r0 r1) (the overload the Python-generator comparison uses)
  void voxel_tables(const std::vector<Real> &rb, int n_sb, std::vector<Real> (&out)[6]) const;

  // thermosphere_exosphere::save (thermosphere_exosphere.cpp:266-300): a text dump of the profile
  virtual void save(const std::string &fname) const;

protected:
  template <class F> Real shell_average(F f, Real r0, Real r1) const;
};

class chamb_diff_1d : public atmosphere_model {
public:
  static const int method_nspmin_nCO2exo = 0;   // thermosphere_exosphere.hpp:27-28
  static const int method_rmax_nCO2rmin = 1;
  Real nH_exo, T_exo, nCO2_exo;                 // n_species_exo / nCO2exo of the reference
  Real rmindiffusion;
  Real n_species_min = 10.0;
  krasnopolsky_temperature temp;
  species_density_parameters species;
  Real lambdac = 0, escape_flux = 0;

  // chamb_diff_1d(n_species_exo, nCO2_exo, &temp, &species) with the default geometry (chamb_diff_1d.cpp:3-14):
  // rmaxx <= 0: rmax = radius where the species falls to n_species_min (thermosphere_exosphere.cpp:57-75)
  chamb_diff_1d(Real nHexo, Real nCO2exo, Real Texo, Real rmaxx = -1);
  // the full constructor (chamb_diff_1d.cpp:15-35)
  chamb_diff_1d(Real rminn, Real rexoo, Real rmaxx_or_nspmin, Real rmindiffusionn, Real n_species_exoo,
                Real nCO2rmin_or_nCO2exoo, const krasnopolsky_temperature &tempp, const species_density_parameters &sp,
                int method = method_nspmin_nCO2exo);

  Real Temp(Real r) const override;
  Real n_species(Real r) const override;
  Real n_absorber(Real r) const override;
  void save(const std::string &fname) const override;

private:
  std::vector<Real> thermo_r, thermo_lnCO2, thermo_lnH;   // ascending radius
  Real n_species_rmindiffusion = 0, nCO2_rmindiffusion = 0;
  Real n_exo(Real r, Real n0, Real mass) const;
  void setup(Real rmaxx_or_nspmin, Real nCO2rmin_or_exo, int method);
  void integrate_thermosphere(int nsteps = 400);
  Real CO2_exobase_from_rmin(Real nCO2rmin) const;        // species_density_parameters::get_CO2_exobase_density
};

// n(r, sza) = n(r) (n0 + nslope sza), CO2 symmetric (chamb_diff_1d_asymmetric.cpp)
class chamb_diff_1d_asymmetric : public chamb_diff_1d {
public:
  Real asymmetry = 1.0, n0 = 1.0, nslope = 0.0;
  using chamb_diff_1d::chamb_diff_1d;
  void set_asymmetry(Real a);
  Real theta_average_factor(Real t0, Real t1) const;
  void voxel_values(Real r0, Real r1, Real t0, Real t1, Real pt_r, Real pt_t, Real (&out)[6]) const override;
  bool sza_dependent() const override { return true; }
};

// exobase temperature linear in SZA from T0 (noon) to T1 (midnight), exobase density A T^-Tpower with the sphere
// average navg; one chamb_diff_1d per SZA node, linear interpolation between nodes (chamb_diff_temp_asymmetric.cpp)
class chamb_diff_temp_asymmetric : public atmosphere_model {
public:
  static const int n_sza = 40;
  Real navg, T0, T1, Tpower, A = 0;
  chamb_diff_temp_asymmetric(const species_density_parameters &sp, Real navgg, Real T00, Real T11, Real nCO2rminn = 2.6e13,
                             Real rexoo = rexo_typical, Real rminn = rMars + 80e5, Real rmaxx = rMars + 50000e5,
                             Real rmindiffusionn = rMars + 80e5, Real T_tropo = 125.0, Real r_tropo = rMars + 90e5,
                             Real shape_parameter = 11.4, Real Tpowerr = 2.5);
  Real T_sza(Real sza) const { return T0 + (T1 - T0) * sza / pi; }
  Real n_species_sza(Real sza) const { return A * std::pow(T_sza(sza), -Tpower); }
  Real Temp(Real r) const override { return atm_sza[0]->Temp(r); }
  Real n_species(Real r) const override { return atm_sza[0]->n_species(r); }
  Real n_absorber(Real r) const override { return atm_sza[0]->n_absorber(r); }
  Real r_from_n_species(Real n) const override { return atm_sza[0]->r_from_n_species(n); }
  Real at(int which, Real r, Real t) const;      // which: 0 species, 1 temperature, 2 absorber
  void voxel_values(Real r0, Real r1, Real t0, Real t1, Real pt_r, Real pt_t, Real (&out)[6]) const override;
  bool sza_dependent() const override { return true; }

private:
  std::vector<std::unique_ptr<chamb_diff_1d>> atm_sza;
};

// linear interpolation of log densities and of temperature on an altitude [km] table; optional Chamberlain exosphere
// above rexo (tabular_atmosphere.cpp)
class tabular_1d : public atmosphere_model {
public:
  bool compute_exosphere = false;
  Real m_species = mH;
  tabular_1d() {}
  tabular_1d(Real rminn, Real rexoo, Real rmaxx, bool compute_exospheree = false);
  void load_log_species_density(const std::vector<double> &alt, const std::vector<double> &log_n_species);
  void load_log_absorber_density(const std::vector<double> &alt, const std::vector<double> &log_n_absorber);
  void load_temperature(const std::vector<double> &alt, const std::vector<double> &temp);
  Real Temp(Real r) const override;
  Real n_species(Real r) const override;
  Real n_absorber(Real r) const override;

private:
  std::vector<double> alt_n, log_n, alt_a, log_a, alt_T, tab_T;
  Real exo_n0 = 0, exo_T = 0, exo_lambdac = 0;
  void check_init();
};

// exobase temperature <-> Jeans parameter / effusion velocity (chamberlain_exosphere.hpp:45-75); the reference
// tabulates 1101 temperatures and interpolates, here the forward maps are evaluated exactly and the inverse maps are
// inverted by bisection on the same [100, 1200] K range
class Temp_converter {
public:
  Temp_converter(double rexoo = rexo_typical, double m_speciess = mH) : rexo(rexoo), m_species(m_speciess) {}
  double lc_from_T_exact(double T) const;
  double eff_from_T_exact(double T) const;
  double lc_from_T(double T) const { return lc_from_T_exact(T); }
  double eff_from_T(double T) const { return eff_from_T_exact(T); }
  double T_from_lc(double lc) const;
  double T_from_eff(double eff) const;

private:
  double rexo, m_species;
};

void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w);   // nodes on [-1, 1]

} // namespace b200rt_host
