// ref_table_atmosphere.hpp -- TEST INFRASTRUCTURE ONLY: shared by oracle/ref_harness*.cpp.
#pragma once
#include <cmath>
#include <vector>
#include "Real.hpp"
#include "atm/atmosphere_base.hpp"
#include "atmo_vec.hpp"

// An `atmosphere` (atm/atmosphere_base.hpp:7-39) whose per-voxel values are
// explicit tables, and whose n_species()/r_from_n_species() pair is rigged so
// that rmethod_log_n_species (grid_spherical...hpp:178-187) reproduces a given
// list of radial boundaries bit for bit: n(rmin)=1, n(rmax)=e^{-(NR-1)} =>
// log-step 1 => target_i = e^{-i} => r_from_n_species returns rb[i].
struct table_atmosphere : atmosphere {
  int nrb;
  std::vector<double> rb;
  std::vector<double> n_avg, n_pt, T_avg, T_pt, nabs_avg, nabs_pt;
  double abs_sigma[2];
  bool spherical;  // touched by observation_fit-style callers; unused here

  table_atmosphere() : atmosphere(0, 0, 0), nrb(0), spherical(true) { abs_sigma[0]=abs_sigma[1]=0; }

  doubReal n_species(const doubReal &r) const override {
    if (r <= rmin) return 1.0;
    return std::exp(-(double)(nrb-1));
  }
  doubReal r_from_n_species(const doubReal &n) const override {
    long i = std::lround(-std::log(n));
    if (i < 0) i = 0;
    if (i > nrb-1) i = nrb-1;
    return rb[i];
  }
  doubReal Temp(const doubReal &) const override { return 0; }
  doubReal n_absorber(const doubReal &) const override { return 0; }

  void n_species_voxel_avg(const atmo_voxel &vox, Real &ret_avg, Real &ret_pt) const {
    ret_avg = n_avg[vox.i_voxel]; ret_pt = n_pt[vox.i_voxel];
  }
  void Temp_voxel_avg(const atmo_voxel &vox, Real &ret_avg, Real &ret_pt) const {
    ret_avg = T_avg[vox.i_voxel]; ret_pt = T_pt[vox.i_voxel];
  }
  void n_absorber_voxel_avg(const atmo_voxel &vox, Real &ret_avg, Real &ret_pt) const {
    ret_avg = nabs_avg[vox.i_voxel]; ret_pt = nabs_pt[vox.i_voxel];
  }
  Real abs_sigma0(const Real &) const { return abs_sigma[0]; }
  Real abs_sigma1(const Real &) const { return abs_sigma[1]; }
};

