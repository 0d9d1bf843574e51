// ref_harness_multiplet.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the reference's OWN multiplet CFR source (emission/multiplet_CFR_emission.hpp with
// O_1026.hpp, H_lyman_multiplet.hpp, H_lyman_multiplet_test.hpp and their trackers), compiled
// unmodified and in place from /root/reference/src by oracle/Makefile, so that
//   * oracle/rt_oracle.c's multiplet restatement can be pinned against it, and
//   * golden fixtures under tests/golden/ can be generated from it.
// No reference source is copied: this file only calls the reference's public API.
//
// kind 0 = O_1026_emission, 1 = H_lyman_multiplet, 2 = H_lyman_singlet (the reference's
// singlet-through-the-multiplet-code consistency check, H_lyman_multiplet_test.hpp).
//
// Build: oracle/_ref/libref_mult_f64.so, libref_mult_f32.so (-DRT_FLOAT).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include <string>
#include <omp.h>

#include "Real.hpp"
#include "constants.hpp"
#include "atm/atmosphere_base.hpp"
#include "atm/atmosphere_average_1d.hpp"
#include "grid/grid_spherical_azimuthally_symmetric.hpp"
#include "RT_grid.hpp"
#include "emission/O_1026.hpp"
#include "emission/H_lyman_multiplet.hpp"
#include "emission/H_lyman_multiplet_test.hpp"
#include "observation.hpp"
#include "ref_table_atmosphere.hpp"

namespace {

// the per-voxel tables of multiplet_CFR_emission are protected; a derived type reads them
template <class E>
struct mult_peek : E {
  static const int NV = E::n_voxels;
  static const int NLOW = E::n_lower, NUP = E::n_upper, NLINES = E::n_lines;
  // [n_lower][2][NV] densities (avg, pt), then T avg, T pt, absorber avg, absorber pt
  void dump_arrays(double *out) const {
    for (int l=0;l<NLOW;l++)
      for (int i=0;i<NV;i++) {
	out[(size_t)(2*l+0)*NV+i] = this->species_density(i,l);
	out[(size_t)(2*l+1)*NV+i] = this->species_density_pt(i,l);
      }
    double *o = out + (size_t) 2*NLOW*NV;
    for (int i=0;i<NV;i++) {
      o[0*NV+i] = this->species_T(i);        o[1*NV+i] = this->species_T_pt(i);
      o[2*NV+i] = this->absorber_density(i); o[3*NV+i] = this->absorber_density_pt(i);
    }
  }
  void dump_K(double *out) const {   // row major, element = voxel*n_upper + state
    const int N = NV*NUP;
    for (int i=0;i<N;i++) for (int j=0;j<N;j++) out[(size_t)i*N+j] = this->influence_matrix(i,j);
  }
  void dump_vectors(double *S0, double *tsp, double *tab, double *S) const {
    for (int i=0;i<NV*NUP;i++) { if (S0) S0[i]=this->singlescat(i); if (S) S[i]=this->sourcefn(i); }
    for (int i=0;i<NV*NLINES;i++) {
      if (tsp) tsp[i]=this->tau_species_single_scattering(i);
      if (tab) tab[i]=this->tau_absorber_single_scattering(i);
    }
  }
  void set_sourcefn(const double *S) { for (int i=0;i<NV*NUP;i++) this->sourcefn(i)=(Real) S[i]; }
  void zero_K() { this->reset_solution(); }
};

// per-kind pieces: solar pumping and the dimensionless line offsets
template <int NV> void set_solar(O_1026_emission<NV> &e, const double *s) { e.set_solar_brightness((Real) s[0]); }
template <int NV> void set_solar(H_lyman_multiplet<NV> &e, const double *s) { e.set_solar_brightness((Real) s[0], (Real) s[1]); }
template <int NV> void set_solar(H_lyman_singlet<NV> &e, const double *s) { e.set_solar_brightness((Real) s[0], (Real) s[1]); }

template <class T> struct waveref;
template <bool b, int NV> struct waveref<O_1026_tracker<b,NV>> {
  static Real get(int) { return O_1026_tracker<b,NV>::doppler_width_wavelength_reference; }
};
template <bool b, int NV> struct waveref<H_lyman_multiplet_tracker<b,NV>> {
  typedef H_lyman_multiplet_tracker<b,NV> T;
  static Real get(int l) { return l < 2 ? T::doppler_width_wavelength_reference_lya : T::doppler_width_wavelength_reference_lyb; }
};
template <bool b, int NV> struct waveref<H_lyman_singlet_tracker<b,NV>> {
  typedef H_lyman_singlet_tracker<b,NV> T;
  static Real get(int l) { return l == 0 ? T::doppler_width_wavelength_reference_lya : T::doppler_width_wavelength_reference_lyb; }
};

struct refm_model {
  virtual ~refm_model() {}
  virtual void dims(int *out) = 0;   // n_vox, n_rays, n_lines, n_multiplets, n_lower, n_upper, n_lambda
  virtual void constants(int *iout, double *dout) = 0;
  virtual int setup(const double *rb, double rexo, int szamethod, int raymethod, const double *solar,
		    const double *vox_in) = 0;
  virtual void get_arrays(double *out) = 0;
  virtual double generate_S() = 0;
  virtual double build_rows(int v0, int v1, int stride, long *n_steps) = 0;
  virtual double solve() = 0;
  virtual void get_K(double *out) = 0;
  virtual void get_vectors(double *S0, double *tsp, double *tab, double *S) = 0;
  virtual void set_sourcefn(const double *S) = 0;
  virtual double brightness(int n, const double *loc, const double *dir, int n_subsamples, double *out) = 0;
  virtual double lineshape(int line, int i_lambda, double T) = 0;
};

template <template <int> class EM, int NR, int NSZA, int NTH, int NPH>
struct refm_impl : refm_model {
  typedef spherical_azimuthally_symmetric_grid<NR,NSZA,NTH,NPH> grid_type;
  static const int NV = grid_type::n_voxels;
  typedef EM<NV> emission_type;
  typedef mult_peek<emission_type> peek_type;
  typedef typename emission_type::brightness_tracker btracker;
  typedef typename emission_type::influence_tracker itracker;
  typedef RT_grid<emission_type, 1, grid_type> RT_type;
  typedef boundary_intersection_stepper<grid_type::n_dimensions, grid_type::n_max_intersections> stepper_type;

  table_atmosphere atm;
  peek_type em;
  emission_type *emp[1];
  RT_type *RT;
  observation<emission_type, 1> *obs;

  refm_impl() : RT(NULL), obs(NULL) { emp[0]=&em; }
  ~refm_impl() { delete obs; delete RT; }

  void dims(int *o) override {
    o[0]=NV; o[1]=grid_type::n_rays; o[2]=btracker::n_lines; o[3]=btracker::n_multiplets;
    o[4]=btracker::n_lower; o[5]=btracker::n_upper; o[6]=btracker::n_lambda;
  }
  // iout[3][n_lines]: multiplet_index, lower_level_index, upper_level_index
  // dout[7][n_lines]: line_sigma_total, line_A, absorber_xsec, upper_state_decay_rate (by upper state; first
  //                   n_upper slots), line_wavelength_offset / doppler_width_wavelength_reference,
  //                   line_shape_normalization at T_ref, weight
  void constants(int *iout, double *dout) override {
    const int NL = btracker::n_lines;
    for (int l=0;l<NL;l++) {
      iout[0*NL+l]=btracker::multiplet_index(l); iout[1*NL+l]=btracker::lower_level_index(l);
      iout[2*NL+l]=btracker::upper_level_index(l);
      dout[0*NL+l]=btracker::line_sigma_total(l); dout[1*NL+l]=btracker::line_A(l);
      dout[2*NL+l]=btracker::absorber_xsec(l);
      dout[3*NL+l]=(l < btracker::n_upper) ? (double) btracker::upper_state_decay_rate(l) : 0.0;
      Real off = btracker::line_wavelength_offset(l) / waveref<btracker>::get(l);
      dout[4*NL+l]=off;
      dout[5*NL+l]=btracker::line_shape_normalization(l, btracker::doppler_width_reference_T);
      dout[6*NL+l]=btracker::weight(l, 0);
    }
  }
  double lineshape(int line, int i_lambda, double T) override {
    return btracker::line_shape_function_normalized(line, i_lambda, (Real) T);
  }

  int setup(const double *rb, double rexo, int szamethod, int raymethod, const double *solar,
	    const double *vox_in) override {
    atm.nrb = NR;
    atm.rb.assign(rb, rb+NR);
    atm.rmin = rb[0]; atm.rexo = rexo; atm.rmax = rb[NR-1];
    atm.n_avg.assign(vox_in+0*NV, vox_in+1*NV);    atm.n_pt.assign(vox_in+1*NV, vox_in+2*NV);
    atm.T_avg.assign(vox_in+2*NV, vox_in+3*NV);    atm.T_pt.assign(vox_in+3*NV, vox_in+4*NV);
    atm.nabs_avg.assign(vox_in+4*NV, vox_in+5*NV); atm.nabs_pt.assign(vox_in+5*NV, vox_in+6*NV);
    delete obs; obs=NULL; delete RT; RT=NULL;
    grid_type *g = new grid_type;
    g->rmethod = 1;                 // rmethod_log_n_species: boundary injection through table_atmosphere
    g->szamethod = szamethod;
    g->raymethod_theta = raymethod;
    g->setup_voxels(atm);
    g->setup_rays();
    em.define("multiplet", atm, &table_atmosphere::n_species_voxel_avg, &table_atmosphere::Temp_voxel_avg,
	      &table_atmosphere::n_absorber_voxel_avg, g->voxels);
    set_solar(em, solar);
    RT = new RT_type(*g, emp);
    delete g;
    obs = new observation<emission_type, 1>(emp);
    return 0;
  }
  void get_arrays(double *out) override { em.dump_arrays(out); }

  double generate_S() override {
    em.zero_K();
    double t0 = omp_get_wtime();
    RT->generate_S();
    return omp_get_wtime()-t0;
  }
  // loop body of RT_grid::generate_S (RT_grid.hpp:160-201) over a strided subset of source voxels, no solve
  double build_rows(int v0, int v1, int stride, long *n_steps) override {
    em.zero_K();
    long steps_total = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel
    {
      itracker *ti = new itracker[1];   // carries n_upper voxel_arrays: keep it off the stack
      ti[0].init();
      itracker (&tref)[1] = *reinterpret_cast<itracker (*)[1]>(ti);
      atmo_vector vec;
#pragma omp for schedule(dynamic,1)
      for (int i_vox=v0; i_vox<v1; i_vox+=stride) {
	for (int i_ray=0; i_ray<grid_type::n_rays; i_ray++) {
	  vec.ptray(RT->grid.voxels[i_vox].pt, RT->grid.rays[i_ray]);
	  emp[0]->reset_tracker(i_vox, tref[0]);
	  RT->voxel_traverse(vec, &RT_type::influence_update, tref);
	  emp[0]->accumulate_influence(i_vox, tref[0]);
	}
	emp[0]->reset_tracker(i_vox, tref[0]);
	RT->get_single_scattering(RT->grid.voxels[i_vox].pt, tref);
      }
      delete [] ti;
    }
    double t = omp_get_wtime()-t0;
    if (n_steps) {
#pragma omp parallel reduction(+:steps_total)
      {
	stepper_type *st = new stepper_type;
	atmo_vector vec;
#pragma omp for schedule(dynamic,1)
	for (int i_vox=v0; i_vox<v1; i_vox+=stride)
	  for (int i_ray=0; i_ray<grid_type::n_rays; i_ray++) {
	    vec.ptray(RT->grid.voxels[i_vox].pt, RT->grid.rays[i_ray]);
	    RT->grid.ray_voxel_intersections(vec, *st);
	    if (st->boundaries.size()>0) steps_total += st->boundaries.size()-1;
	  }
	delete st;
      }
      *n_steps = steps_total;
    }
    return t;
  }
  double solve() override { double t0 = omp_get_wtime(); RT->solve(); return omp_get_wtime()-t0; }
  void get_K(double *out) override { em.dump_K(out); }
  void get_vectors(double *S0, double *tsp, double *tab, double *S) override { em.dump_vectors(S0,tsp,tab,S); }
  void set_sourcefn(const double *S) override { em.set_sourcefn(S); }

  // out[(3*n_lines + n_lower)][n]: brightness[line], tau_species_final[line], tau_absorber_final[line],
  // species_col_dens[lower]
  double brightness(int n, const double *loc, const double *dir, int n_subsamples, double *out) override {
    std::vector<std::vector<Real>> L(n, std::vector<Real>(3)), D(n, std::vector<Real>(3));
    for (int i=0;i<n;i++) for (int k=0;k<3;k++) { L[i][k]=(Real) loc[3*i+k]; D[i][k]=(Real) dir[3*i+k]; }
    obs->add_MSO_observation(L, D);
    std::streambuf *cb = std::cout.rdbuf();
    std::cout.rdbuf(NULL);
    double t0 = omp_get_wtime();
    RT->brightness(*obs, n_subsamples);
    double t = omp_get_wtime()-t0;
    std::cout.rdbuf(cb);
    const int NL = btracker::n_lines, NLOW = btracker::n_lower;
    for (int i=0;i<n;i++) {
      const btracker &b = obs->los[0][i];
      for (int l=0;l<NL;l++) {
	out[(size_t)(0*NL+l)*n+i]=b.brightness[l];
	out[(size_t)(1*NL+l)*n+i]=b.tau_species_final[l];
	out[(size_t)(2*NL+l)*n+i]=b.tau_absorber_final[l];
      }
      for (int l=0;l<NLOW;l++) out[(size_t)(3*NL+l)*n+i]=b.species_col_dens[l];
    }
    return t;
  }
};

template <int NR, int NSZA, int NTH, int NPH>
refm_model* make_kind(int kind) {
  if (kind==0) return new refm_impl<O_1026_emission,NR,NSZA,NTH,NPH>;
  if (kind==1) return new refm_impl<H_lyman_multiplet,NR,NSZA,NTH,NPH>;
  if (kind==2) return new refm_impl<H_lyman_singlet,NR,NSZA,NTH,NPH>;
  return NULL;
}

} // namespace

// shapes: two small test grids and the reference default 40x20x7x12 (observation_fit.hpp:44-47)
#define REFM_SHAPES X(8,6,4,4) X(12,8,5,6) X(40,20,7,12)

extern "C" {
void* refm_create(int kind, int NR, int NSZA, int NTH, int NPH) {
#define X(a,b,c,d) if (NR==a && NSZA==b && NTH==c && NPH==d) return make_kind<a,b,c,d>(kind);
  REFM_SHAPES
#undef X
  return NULL;
}
void refm_destroy(void *h) { delete static_cast<refm_model*>(h); }
void refm_dims(void *h, int *out) { static_cast<refm_model*>(h)->dims(out); }
void refm_constants(void *h, int *iout, double *dout) { static_cast<refm_model*>(h)->constants(iout, dout); }
double refm_lineshape(void *h, int line, int i_lambda, double T) { return static_cast<refm_model*>(h)->lineshape(line, i_lambda, T); }
int refm_setup(void *h, const double *rb, double rexo, int szamethod, int raymethod, const double *solar,
	       const double *vox_in) {
  return static_cast<refm_model*>(h)->setup(rb, rexo, szamethod, raymethod, solar, vox_in);
}
void refm_get_arrays(void *h, double *out) { static_cast<refm_model*>(h)->get_arrays(out); }
double refm_generate_S(void *h) { return static_cast<refm_model*>(h)->generate_S(); }
double refm_build_rows(void *h, int v0, int v1, int stride, long *n_steps) { return static_cast<refm_model*>(h)->build_rows(v0, v1, stride, n_steps); }
double refm_solve(void *h) { return static_cast<refm_model*>(h)->solve(); }
void refm_get_K(void *h, double *out) { static_cast<refm_model*>(h)->get_K(out); }
void refm_get_vectors(void *h, double *S0, double *tsp, double *tab, double *S) { static_cast<refm_model*>(h)->get_vectors(S0, tsp, tab, S); }
void refm_set_sourcefn(void *h, const double *S) { static_cast<refm_model*>(h)->set_sourcefn(S); }
double refm_brightness(void *h, int n, const double *loc, const double *dir, int n_subsamples, double *out) {
  return static_cast<refm_model*>(h)->brightness(n, loc, dir, n_subsamples, out);
}
}
