// atmosphere.hpp -- Boost-free restatement of the reference's default atmosphere chamb_diff_1d
// (src/atm/chamb_diff_1d.*, thermosphere_exosphere.*, chamberlain_exosphere.*, temperature.*,
// species_density_parameters.*, atmosphere_average_1d.*).
//
// This is the host-side INPUT GENERATOR of observation_fit::generate_source_function(nH, T)
// (observation_fit.cpp:122-135): it runs once per parameter set and feeds per-voxel arrays to the
// device path.  The reference builds it on Boost (gamma_p, odeint, cubic B-splines, root bracketing),
// which is not available; parity of the hot path is evaluated at the voxel-array boundary, so this
// class only has to be a faithful physical restatement, and it is numerically the same algorithm as
// the Python generator the tests use (3d_planetary_rt_model_b200/synth.py) so that the facade and the
// oracle pipeline can be compared end to end:
//   temperature  Krasnopolsky profile                         temperature.cpp:23-48
//   exosphere    Chamberlain, P(3/2,x) = erf(sqrt x) - 2 sqrt(x/pi) e^-x     chamberlain_exosphere.cpp:22-59
//   thermosphere diffusive H in CO2, classic RK4 from the exobase down       species_density_parameters.cpp:83-156
//   averages     shell averages by Gauss-Legendre quadrature in log r        atmosphere_average_1d.cpp:141-157
#pragma once
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
};

class chamb_diff_1d : public H_cross_sections {
public:
  Real nH_exo, T_exo, nCO2_exo;
  Real rmin = rMars + 80e5, rexo = rMars + 200e5, rmax = 0;
  Real n_species_min = 10.0;
  Real T_tropo = 125.0, r_tropo = rMars + 90e5, shape = 11.4;     // krasnopolsky_temperature, temperature.hpp:39-43
  Real lambdac = 0, escape_flux = 0;

  // rmaxx <= 0: rmax = radius where n_H falls to n_species_min (thermosphere_exosphere.cpp:57-75)
  chamb_diff_1d(Real nHexo, Real nCO2exo, Real Texo, Real rmaxx = -1);

  Real Temp(Real r) const;
  Real n_species(Real r) const;
  Real n_absorber(Real r) const;
  Real r_from_n_species(Real n) const;

  // radial boundaries: rmethod 0 = altitude (coordinate_generation.hpp:57-87), 1 = log n_species
  // (grid_spherical_azimuthally_symmetric.hpp:178-187)
  std::vector<Real> radial_boundaries(int n_rb, int rmethod) const;

  // per-voxel inputs of singlet_CFR::define (singlet_CFR.hpp:419-492): [6][n_vox] = n_avg, n_pt, T_avg, T_pt,
  // nabs_avg, nabs_pt with voxel id = ir*(n_sb-1)+isza (1-D atmosphere: the same for every isza)
  void voxel_tables(const std::vector<Real> &rb, int n_sb, std::vector<Real> (&out)[6]) const;

private:
  std::vector<Real> thermo_r, thermo_lnCO2, thermo_lnH;   // ascending radius
  Real Tprime(Real r) const;
  Real n_exo(Real r) const;
  void integrate_thermosphere(int nsteps = 400);
  template <class F> Real shell_average(F f, Real r0, Real r1) const;
};

void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w);   // nodes on [-1, 1]

} // namespace b200rt_host
