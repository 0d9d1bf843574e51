// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the reference's OWN hot-path source, compiled unmodified and in place
// from /root/reference/src (see oracle/Makefile), so that
//   * oracle/rt_oracle.c (our CPU restatement) can be pinned against it,
//   * golden fixtures under tests/golden/ can be generated from it, and
//   * bench.py --impl reference can time the reference CPU (OpenMP) path.
// No reference source is copied: this file only *calls* the reference's public
// API (RT_grid, spherical_azimuthally_symmetric_grid, singlet_CFR, observation).
//
// The reference's grid sizes are template ints (observation_fit.hpp:44-47), so a
// fixed list of shapes is instantiated below and selected at run time.
//
// Build: one shared object per precision (Real = double, or float with
// -DRT_FLOAT, Real.hpp:9-27):  oracle/_ref/libref_f64.so, oracle/_ref/libref_f32.so
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <omp.h>

#include "Real.hpp"
#include "constants.hpp"
#include "atm/atmosphere_base.hpp"
#include "atm/atmosphere_average_1d.hpp"  // completes the type named at grid_spherical...hpp:270
#include "grid/grid_spherical_azimuthally_symmetric.hpp"
#include "grid/grid_plane_parallel.hpp"
#include "RT_grid.hpp"
#include "emission/singlet_CFR.hpp"
#include "observation.hpp"
#include "ref_table_atmosphere.hpp"
#ifdef RT_B200
#include "RT_b200.hpp"   // integration/: the reference-side binding of libb200rt.so (defines the RT_grid *_gpu members)
#endif

namespace {

// singlet_CFR keeps its per-voxel tables protected; a derived type reads them.
template <int NV>
struct singlet_peek : singlet_CFR<NV> {
  typedef singlet_CFR<NV> base;
  void dump_arrays(double *out) const {  // [10][NV]
    for (int i=0;i<NV;i++) {
      out[0*NV+i]=base::species_T_ratio(i);   out[1*NV+i]=base::species_T_ratio_pt(i);
      out[2*NV+i]=base::species_density(i);   out[3*NV+i]=base::species_density_pt(i);
      out[4*NV+i]=base::dtau_species(i);      out[5*NV+i]=base::dtau_species_pt(i);
      out[6*NV+i]=base::dtau_absorber(i);     out[7*NV+i]=base::dtau_absorber_pt(i);
      out[8*NV+i]=base::abs(i);               out[9*NV+i]=base::abs_pt(i);
    }
  }
  void dump_K(double *out) const { // row major NVxNV
    for (int i=0;i<NV;i++) for (int j=0;j<NV;j++) out[(size_t)i*NV+j]=base::influence_matrix(i,j);
  }
  void dump_vectors(double *S0, double *tsp, double *tab, double *S) const {
    for (int i=0;i<NV;i++) {
      if (S0)  S0[i]=base::singlescat(i);
      if (tsp) tsp[i]=base::tau_species_single_scattering(i);
      if (tab) tab[i]=base::tau_absorber_single_scattering(i);
      if (S)   S[i]=base::sourcefn(i);
    }
  }
  void set_sourcefn(const double *S) { for (int i=0;i<NV;i++) base::sourcefn(i)=(Real) S[i]; }
  void zero_K() { base::reset_solution(); }
};

struct ref_model {
  virtual ~ref_model() {}
  virtual int setup(const double *rb, double rexo, int rmethod, int szamethod, int raymethod,
		    int n_em, const double *em_scalars, const double *abs_sigma, const double *vox_in) = 0;
  virtual void get_grid(double *sza_b, double *pts_r, double *pts_sza, double *ray_theta, double *ray_phi,
			double *ray_domega, double *rad_b) = 0;
  virtual void get_arrays(int e, double *out) = 0;
  virtual long traverse_voxel_rays(int v0, int v1, long cap, int *len, int *exits_bottom,
				   int *entering, double *distance) = 0;
  virtual long traverse_los(int n, const double *loc, const double *dir, long cap, int *len, int *exits_bottom,
			    int *entering, double *distance, double *rayscal) = 0;
  virtual double generate_S() = 0;
  virtual double build_rows(int v0, int v1, int stride, long *n_steps) = 0;
  virtual double solve() = 0;
  virtual void get_K(int e, double *out) = 0;
  virtual void get_vectors(int e, double *S0, double *tsp, double *tab, double *S) = 0;
  virtual void set_sourcefn(int e, const double *S) = 0;
  virtual double brightness(int n, const double *loc, const double *dir, int n_subsamples, double *out) = 0;
  virtual int n_voxels() = 0;
  virtual int n_rays() = 0;
  // the reference's own ASCII writers (RT_grid.hpp:221-230): 0 = done, -1 = not available on this model
  // RT_grid::generate_S_gpu / brightness_gpu through integration/RT_b200.hpp (-DRT_B200 builds only): -1 = not built in
  virtual int generate_S_gpu() { return -1; }
  virtual int brightness_gpu(int, const double *, const double *, int, double *) { return -1; }
  virtual int influence_to_host() { return -1; }
  virtual int save_S(const char *) { return -1; }
  virtual int save_influence(const char *) { return -1; }
};

template <int NR, int NSZA, int NTH, int NPH, int NEM>
struct ref_model_impl : ref_model {
  typedef spherical_azimuthally_symmetric_grid<NR,NSZA,NTH,NPH> grid_type;
  static const int NV = grid_type::n_voxels;
  typedef singlet_CFR<NV> emission_type;
  typedef singlet_peek<NV> peek_type;
  typedef RT_grid<emission_type, NEM, grid_type> RT_type;
  typedef boundary_intersection_stepper<grid_type::n_dimensions, grid_type::n_max_intersections> stepper_type;

  table_atmosphere atm;
  peek_type em[NEM];
  emission_type *emp[NEM];
  RT_type *RT;
  observation<emission_type, NEM> *obs;

  ref_model_impl() : RT(NULL), obs(NULL) {
    for (int e=0;e<NEM;e++) emp[e]=&em[e];
  }
  ~ref_model_impl() { delete obs; delete RT; }
  int n_voxels() override { return NV; }
  int n_rays() override { return grid_type::n_rays; }

  int setup(const double *rb, double rexo, int rmethod, int szamethod, int raymethod,
	    int n_em, const double *em_scalars, const double *abs_sigma, const double *vox_in) override {
    if (n_em != NEM) return -2;
    atm.nrb = NR;
    atm.rb.assign(rb, rb+NR);
    atm.rmin = rb[0]; atm.rexo = rexo; atm.rmax = rb[NR-1];
    atm.n_avg.assign(vox_in+0*NV, vox_in+1*NV);    atm.n_pt.assign(vox_in+1*NV, vox_in+2*NV);
    atm.T_avg.assign(vox_in+2*NV, vox_in+3*NV);    atm.T_pt.assign(vox_in+3*NV, vox_in+4*NV);
    atm.nabs_avg.assign(vox_in+4*NV, vox_in+5*NV); atm.nabs_pt.assign(vox_in+5*NV, vox_in+6*NV);
    for (int e=0;e<NEM;e++) atm.abs_sigma[e]=abs_sigma[e];

    delete obs; obs=NULL; delete RT; RT=NULL;
    grid_type *g = new grid_type;   // the grid struct is large (B: 680 kB): keep it off the stack
    g->rmethod = rmethod;           // 0 = rmethod_altitude, 1 = rmethod_log_n_species (boundary injection)
    g->szamethod = szamethod;
    g->raymethod_theta = raymethod;
    g->setup_voxels(atm);
    g->setup_rays();
    for (int e=0;e<NEM;e++) {
      const double *s = em_scalars+4*e;
      char name[32]; snprintf(name, sizeof(name), "emission %d", e);
      em[e].define(name, (Real) s[0], (Real) s[1], (Real) s[2], atm,
		   &table_atmosphere::n_species_voxel_avg, &table_atmosphere::Temp_voxel_avg,
		   &table_atmosphere::n_absorber_voxel_avg,
		   e==0 ? &table_atmosphere::abs_sigma0 : &table_atmosphere::abs_sigma1,
		   g->voxels);
      em[e].set_emission_g_factor((Real) s[3]);
    }
    RT = new RT_type(*g, emp);
    delete g;
    obs = new observation<emission_type, NEM>(emp);
    return 0;
  }

  void get_grid(double *sza_b, double *pts_r, double *pts_sza, double *ray_theta, double *ray_phi,
		double *ray_domega, double *rad_b) override {
    const grid_type &g = RT->grid;
    for (int i=0;i<NSZA;i++) sza_b[i]=g.sza_boundaries[i];
    for (int i=0;i<NR-1;i++) pts_r[i]=g.pts_radii[i];
    for (int i=0;i<NSZA-1;i++) pts_sza[i]=g.pts_sza[i];
    for (int i=0;i<NTH;i++) ray_theta[i]=g.ray_theta[i];
    for (int i=0;i<NPH;i++) ray_phi[i]=g.ray_phi[i];
    for (int i=0;i<NTH*NPH;i++) ray_domega[i]=g.rays[i].domega;
    for (int i=0;i<NR;i++) rad_b[i]=g.radial_boundaries[i];
  }
  void get_arrays(int e, double *out) override { em[e].dump_arrays(out); }
  int save_S(const char *fname) override { RT->save_S(fname); return 0; }
  int save_influence(const char *fname) override { RT->save_influence(fname); return 0; }
#ifdef RT_B200
  int generate_S_gpu() override { RT->generate_S_gpu(); return 0; }
  int influence_to_host() override { RT->emissions_influence_to_host(); return 0; }
  int brightness_gpu(int n, const double *loc, const double *dir, int n_subsamples, double *out) override {
    std::vector<std::vector<Real>> L(n, std::vector<Real>(3)), D(n, std::vector<Real>(3));
    for (int i=0;i<n;i++) for (int k=0;k<3;k++) { L[i][k]=(Real) loc[3*i+k]; D[i][k]=(Real) dir[3*i+k]; }
    obs->add_MSO_observation(L, D);
    RT->brightness_gpu(*obs, n_subsamples);
    for (int e=0;e<NEM;e++)
      for (int i=0;i<n;i++) {
	out[((size_t)e*4+0)*n+i]=obs->los[e][i].brightness;
	out[((size_t)e*4+1)*n+i]=obs->los[e][i].tau_species_final;
	out[((size_t)e*4+2)*n+i]=obs->los[e][i].tau_absorber_final;
	out[((size_t)e*4+3)*n+i]=obs->los[e][i].species_col_dens;
      }
    return 0;
  }
#endif

  long dump_stepper(const stepper_type &st, long pos, long cap, int *len, int *exits_bottom,
		    int *entering, double *distance) {
    const int n = st.boundaries.size();
    *len = n;
    *exits_bottom = (n>0 && st.exits_bottom) ? 1 : 0;
    for (int k=0;k<n;k++) {
      if (pos+k >= cap) return -1;
      entering[pos+k] = st.boundaries[k].entering;
      distance[pos+k] = st.boundaries[k].distance;
    }
    return pos+n;
  }

  // boundary lists of voxel-origin rays (what generate_S marches), RT_grid.hpp:172-181
  long traverse_voxel_rays(int v0, int v1, long cap, int *len, int *exits_bottom,
			   int *entering, double *distance) override {
    long pos=0;
    stepper_type *st = new stepper_type;
    for (int iv=v0; iv<v1; iv++)
      for (int ir=0; ir<grid_type::n_rays; ir++) {
	atmo_vector vec;
	vec.ptray(RT->grid.voxels[iv].pt, RT->grid.rays[ir]);
	RT->grid.ray_voxel_intersections(vec, *st);
	long idx = (long)(iv-v0)*grid_type::n_rays+ir;
	pos = dump_stepper(*st, pos, cap, len+idx, exits_bottom+idx, entering, distance);
	if (pos<0) { delete st; return -1; }
      }
    delete st;
    return pos;
  }

  // boundary lists for observer lines of sight given in MSO coordinates
  // (observation.hpp:46-65 -> atmo_vec.cpp:256-290); rayscal[n][6] returns the
  // ray scalars the traversal depends on: r, z, t, cost, line_z, (unused)
  long traverse_los(int n, const double *loc, const double *dir, long cap, int *len, int *exits_bottom,
		    int *entering, double *distance, double *rayscal) override {
    std::vector<std::vector<Real>> L(n, std::vector<Real>(3)), D(n, std::vector<Real>(3));
    for (int i=0;i<n;i++) for (int k=0;k<3;k++) { L[i][k]=(Real) loc[3*i+k]; D[i][k]=(Real) dir[3*i+k]; }
    obs->add_MSO_observation(L, D);
    long pos=0;
    stepper_type *st = new stepper_type;
    for (int i=0;i<n;i++) {
      atmo_vector vec = obs->get_vec(i);
      if (rayscal) {
	rayscal[6*i+0]=vec.pt.r; rayscal[6*i+1]=vec.pt.z; rayscal[6*i+2]=vec.pt.t;
	rayscal[6*i+3]=vec.ray.cost; rayscal[6*i+4]=vec.line_z; rayscal[6*i+5]=vec.line_x;
      }
      RT->grid.ray_voxel_intersections(vec, *st);
      pos = dump_stepper(*st, pos, cap, len+i, exits_bottom+i, entering, distance);
      if (pos<0) { delete st; return -1; }
    }
    delete st;
    return pos;
  }

  // the reference's own driver, untouched: influence build + solve (RT_grid.hpp:150-218)
  double generate_S() override {
    for (int e=0;e<NEM;e++) em[e].zero_K();
    double t0 = omp_get_wtime();
    RT->generate_S();
    return omp_get_wtime()-t0;
  }

  // Same loop body as RT_grid::generate_S (RT_grid.hpp:160-201), calling the
  // same reference member functions, but over a strided subset of source voxels
  // and without the solve, so that large grids can be sampled and the two phases
  // timed separately.  Returns wall seconds; *n_steps = ray-voxel steps executed.
  double build_rows(int v0, int v1, int stride, long *n_steps) override {
    for (int e=0;e<NEM;e++) em[e].zero_K();
    long steps_total = 0;
    double t0 = omp_get_wtime();
#pragma omp parallel
    {
      typename emission_type::influence_tracker ti[NEM];
      for (int e=0;e<NEM;e++) ti[e].init();
      atmo_vector vec;
#pragma omp for schedule(dynamic,1)
      for (int i_vox=v0; i_vox<v1; i_vox+=stride) {
	for (int i_ray=0; i_ray<grid_type::n_rays; i_ray++) {
	  vec.ptray(RT->grid.voxels[i_vox].pt, RT->grid.rays[i_ray]);
	  for (int e=0;e<NEM;e++) emp[e]->reset_tracker(i_vox, ti[e]);
	  RT->voxel_traverse(vec, &RT_type::influence_update, ti);
	  for (int e=0;e<NEM;e++) emp[e]->accumulate_influence(i_vox, ti[e]);
	}
	for (int e=0;e<NEM;e++) emp[e]->reset_tracker(i_vox, ti[e]);
	RT->get_single_scattering(RT->grid.voxels[i_vox].pt, ti);
      }
    }
    double t = omp_get_wtime()-t0;
    if (n_steps) {
      // step count, computed after the timed region
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

  double solve() override {
    double t0 = omp_get_wtime();
    RT->solve();
    return omp_get_wtime()-t0;
  }

  void get_K(int e, double *out) override { em[e].dump_K(out); }
  void get_vectors(int e, double *S0, double *tsp, double *tab, double *S) override { em[e].dump_vectors(S0,tsp,tab,S); }
  void set_sourcefn(int e, const double *S) override { em[e].set_sourcefn(S); }

  // out[NEM][4][n]: brightness, tau_species_final, tau_absorber_final, species_col_dens
  // (the four quantities observation_fit.cpp:491-559 reads back)
  double brightness(int n, const double *loc, const double *dir, int n_subsamples, double *out) override {
    std::vector<std::vector<Real>> L(n, std::vector<Real>(3)), D(n, std::vector<Real>(3));
    for (int i=0;i<n;i++) for (int k=0;k<3;k++) { L[i][k]=(Real) loc[3*i+k]; D[i][k]=(Real) dir[3*i+k]; }
    obs->add_MSO_observation(L, D);
    // RT_grid::brightness(obs) prints its (CPU-time) clock; silence it
    std::streambuf *cb = std::cout.rdbuf();
    std::cout.rdbuf(NULL);
    double t0 = omp_get_wtime();
    RT->brightness(*obs, n_subsamples);
    double t = omp_get_wtime()-t0;
    std::cout.rdbuf(cb);
    for (int e=0;e<NEM;e++)
      for (int i=0;i<n;i++) {
	out[((size_t)e*4+0)*n+i]=obs->los[e][i].brightness;
	out[((size_t)e*4+1)*n+i]=obs->los[e][i].tau_species_final;
	out[((size_t)e*4+2)*n+i]=obs->los[e][i].tau_absorber_final;
	out[((size_t)e*4+3)*n+i]=obs->los[e][i].species_col_dens;
      }
    return t;
  }
};

// ---- plane-parallel grid (grid/grid_plane_parallel.hpp): source function only -- the reference has no
// interp_weights on this grid (:304-311), so there is no line-of-sight brightness to pin
template <int NR, int NTH, int NEM>
struct ref_model_pp : ref_model {
  typedef plane_parallel_grid<NR,NTH> grid_type;
  static const int NV = grid_type::n_voxels;
  typedef singlet_CFR<NV> emission_type;
  typedef singlet_peek<NV> peek_type;
  typedef RT_grid<emission_type, NEM, grid_type> RT_type;
  typedef boundary_intersection_stepper<grid_type::n_dimensions, grid_type::n_max_intersections> stepper_type;

  table_atmosphere atm;
  peek_type em[NEM];
  emission_type *emp[NEM];
  RT_type *RT;
  ref_model_pp() : RT(NULL) { for (int e=0;e<NEM;e++) emp[e]=&em[e]; }
  ~ref_model_pp() { delete RT; }
  int n_voxels() override { return NV; }
  int n_rays() override { return grid_type::n_rays; }

  int setup(const double *rb, double rexo, int rmethod, int, int, int n_em, const double *em_scalars,
	    const double *abs_sigma, const double *vox_in) override {
    if (n_em != NEM) return -2;
    atm.nrb = NR;
    atm.rb.assign(rb, rb+NR);
    atm.rmin = rb[0]; atm.rexo = rexo; atm.rmax = rb[NR-1];
    atm.n_avg.assign(vox_in+0*NV, vox_in+1*NV);    atm.n_pt.assign(vox_in+1*NV, vox_in+2*NV);
    atm.T_avg.assign(vox_in+2*NV, vox_in+3*NV);    atm.T_pt.assign(vox_in+3*NV, vox_in+4*NV);
    atm.nabs_avg.assign(vox_in+4*NV, vox_in+5*NV); atm.nabs_pt.assign(vox_in+5*NV, vox_in+6*NV);
    for (int e=0;e<NEM;e++) atm.abs_sigma[e]=abs_sigma[e];
    delete RT; RT=NULL;
    grid_type *g = new grid_type;
    g->rmethod = rmethod;           // 1 = rmethod_log_n_species (boundary injection through table_atmosphere)
    g->setup_voxels(atm);
    g->setup_rays();
    for (int e=0;e<NEM;e++) {
      const double *s = em_scalars+4*e;
      char name[32]; snprintf(name, sizeof(name), "emission %d", e);
      em[e].define(name, (Real) s[0], (Real) s[1], (Real) s[2], atm,
		   &table_atmosphere::n_species_voxel_avg, &table_atmosphere::Temp_voxel_avg,
		   &table_atmosphere::n_absorber_voxel_avg,
		   e==0 ? &table_atmosphere::abs_sigma0 : &table_atmosphere::abs_sigma1,
		   g->voxels);
      em[e].set_emission_g_factor((Real) s[3]);
    }
    RT = new RT_type(*g, emp);
    delete g;
    return 0;
  }
  // sza_b / pts_sza / ray_phi are not filled (1-D grid)
  void get_grid(double *, double *pts_r, double *, double *ray_theta, double *, double *ray_domega, double *rad_b) override {
    const grid_type &g = RT->grid;
    for (int i=0;i<NR-1;i++) pts_r[i]=g.pts_radii[i];
    for (int i=0;i<NTH;i++) { ray_theta[i]=g.rays[i].t; ray_domega[i]=g.rays[i].domega; }
    for (int i=0;i<NR;i++) rad_b[i]=g.radial_boundaries[i];
  }
  void get_arrays(int e, double *out) override { em[e].dump_arrays(out); }
  long traverse_voxel_rays(int v0, int v1, long cap, int *len, int *exits_bottom, int *entering, double *distance) override {
    long pos=0;
    stepper_type *st = new stepper_type;
    for (int iv=v0; iv<v1; iv++)
      for (int ir=0; ir<grid_type::n_rays; ir++) {
	atmo_vector vec;
	vec.ptray(RT->grid.voxels[iv].pt, RT->grid.rays[ir]);
	RT->grid.ray_voxel_intersections(vec, *st);
	long idx = (long)(iv-v0)*grid_type::n_rays+ir;
	const int n = st->boundaries.size();
	len[idx] = n;
	exits_bottom[idx] = (n>0 && st->exits_bottom) ? 1 : 0;
	for (int k=0;k<n;k++) {
	  if (pos+k >= cap) { delete st; return -1; }
	  entering[pos+k] = st->boundaries[k].entering;
	  distance[pos+k] = st->boundaries[k].distance;
	}
	pos += n;
      }
    delete st;
    return pos;
  }
  long traverse_los(int, const double *, const double *, long, int *, int *, int *, double *, double *) override { return -1; }
  double generate_S() override {
    for (int e=0;e<NEM;e++) em[e].zero_K();
    double t0 = omp_get_wtime();
    RT->generate_S();
    return omp_get_wtime()-t0;
  }
  double build_rows(int v0, int v1, int stride, long *n_steps) override {
    for (int e=0;e<NEM;e++) em[e].zero_K();
    long steps_total = 0;
    double t0 = omp_get_wtime();
    {
      typename emission_type::influence_tracker ti[NEM];
      for (int e=0;e<NEM;e++) ti[e].init();
      atmo_vector vec;
      stepper_type *st = new stepper_type;
      for (int i_vox=v0; i_vox<v1; i_vox+=stride) {
	for (int i_ray=0; i_ray<grid_type::n_rays; i_ray++) {
	  vec.ptray(RT->grid.voxels[i_vox].pt, RT->grid.rays[i_ray]);
	  for (int e=0;e<NEM;e++) emp[e]->reset_tracker(i_vox, ti[e]);
	  RT->voxel_traverse(vec, &RT_type::influence_update, ti);
	  for (int e=0;e<NEM;e++) emp[e]->accumulate_influence(i_vox, ti[e]);
	  RT->grid.ray_voxel_intersections(vec, *st);
	  if (st->boundaries.size()>0) steps_total += st->boundaries.size()-1;
	}
	for (int e=0;e<NEM;e++) emp[e]->reset_tracker(i_vox, ti[e]);
	RT->get_single_scattering(RT->grid.voxels[i_vox].pt, ti);
      }
      delete st;
    }
    if (n_steps) *n_steps = steps_total;
    return omp_get_wtime()-t0;
  }
  double solve() override { double t0 = omp_get_wtime(); RT->solve(); return omp_get_wtime()-t0; }
  void get_K(int e, double *out) override { em[e].dump_K(out); }
  void get_vectors(int e, double *S0, double *tsp, double *tab, double *S) override { em[e].dump_vectors(S0,tsp,tab,S); }
  void set_sourcefn(int e, const double *S) override { em[e].set_sourcefn(S); }
  double brightness(int, const double *, const double *, int, double *) override { return -1; }
};

template <int NR, int NTH>
ref_model* make_model_pp(int n_em) {
  if (n_em==1) return new ref_model_pp<NR,NTH,1>;
  if (n_em==2) return new ref_model_pp<NR,NTH,2>;
  return NULL;
}

template <int NR, int NSZA, int NTH, int NPH>
ref_model* make_model(int n_em) {
  if (n_em==1) return new ref_model_impl<NR,NSZA,NTH,NPH,1>;
  if (n_em==2) return new ref_model_impl<NR,NSZA,NTH,NPH,2>;
  return NULL;
}

} // namespace

// shapes instantiated: D = reference default (generate_source_function.cpp:95-98,
// observation_fit.hpp:44-47); B = BASELINE.json configs[1]; the rest are small
// test grids.  Extra shapes can be added with -DREF_EXTRA_SHAPES="X(a,b,c,d)".
#ifndef REF_EXTRA_SHAPES
#define REF_EXTRA_SHAPES
#endif
#define REF_SHAPES X(40,20,7,12) X(100,60,24,16) X(12,8,5,6) X(20,12,6,8) X(8,6,4,4) REF_EXTRA_SHAPES

extern "C" {

int ref_real_bytes() { return (int) sizeof(Real); }

void* ref_create(int NR, int NSZA, int NTH, int NPH, int n_em) {
#define X(a,b,c,d) if (NR==a && NSZA==b && NTH==c && NPH==d) return make_model<a,b,c,d>(n_em);
  REF_SHAPES
#undef X
  return NULL;
}
// plane-parallel shapes: observation_fit's <40, 7> (observation_fit.hpp:44-50), generate_source_function.cpp's
// <40, 6> (:85-93) and two small test grids
#define REF_PP_SHAPES Y(40,7) Y(40,6) Y(12,5) Y(8,4)
void* ref_create_pp(int NR, int NTH, int n_em) {
#define Y(a,c) if (NR==a && NTH==c) return make_model_pp<a,c>(n_em);
  REF_PP_SHAPES
#undef Y
  return NULL;
}
void ref_destroy(void *h) { delete static_cast<ref_model*>(h); }
int ref_n_voxels(void *h) { return static_cast<ref_model*>(h)->n_voxels(); }
int ref_n_rays(void *h) { return static_cast<ref_model*>(h)->n_rays(); }

int ref_setup(void *h, const double *rb, double rexo, int rmethod, int szamethod, int raymethod,
	      int n_em, const double *em_scalars, const double *abs_sigma, const double *vox_in) {
  return static_cast<ref_model*>(h)->setup(rb, rexo, rmethod, szamethod, raymethod, n_em, em_scalars, abs_sigma, vox_in);
}
void ref_get_grid(void *h, double *sza_b, double *pts_r, double *pts_sza, double *ray_theta, double *ray_phi,
		  double *ray_domega, double *rad_b) {
  static_cast<ref_model*>(h)->get_grid(sza_b, pts_r, pts_sza, ray_theta, ray_phi, ray_domega, rad_b);
}
void ref_get_arrays(void *h, int e, double *out) { static_cast<ref_model*>(h)->get_arrays(e, out); }
long ref_traverse_voxel_rays(void *h, int v0, int v1, long cap, int *len, int *exits_bottom, int *entering, double *distance) {
  return static_cast<ref_model*>(h)->traverse_voxel_rays(v0, v1, cap, len, exits_bottom, entering, distance);
}
long ref_traverse_los(void *h, int n, const double *loc, const double *dir, long cap, int *len, int *exits_bottom,
		      int *entering, double *distance, double *rayscal) {
  return static_cast<ref_model*>(h)->traverse_los(n, loc, dir, cap, len, exits_bottom, entering, distance, rayscal);
}
double ref_generate_S(void *h) { return static_cast<ref_model*>(h)->generate_S(); }
double ref_build_rows(void *h, int v0, int v1, int stride, long *n_steps) {
  return static_cast<ref_model*>(h)->build_rows(v0, v1, stride, n_steps);
}
double ref_solve(void *h) { return static_cast<ref_model*>(h)->solve(); }
void ref_get_K(void *h, int e, double *out) { static_cast<ref_model*>(h)->get_K(e, out); }
void ref_get_vectors(void *h, int e, double *S0, double *tsp, double *tab, double *S) {
  static_cast<ref_model*>(h)->get_vectors(e, S0, tsp, tab, S);
}
void ref_set_sourcefn(void *h, int e, const double *S) { static_cast<ref_model*>(h)->set_sourcefn(e, S); }
double ref_brightness(void *h, int n, const double *loc, const double *dir, int n_subsamples, double *out) {
  return static_cast<ref_model*>(h)->brightness(n, loc, dir, n_subsamples, out);
}
int ref_generate_S_gpu(void *h) { try { return static_cast<ref_model*>(h)->generate_S_gpu(); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return -3; } }
int ref_influence_to_host(void *h) { try { return static_cast<ref_model*>(h)->influence_to_host(); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return -3; } }
int ref_brightness_gpu(void *h, int n, const double *loc, const double *dir, int n_subsamples, double *out) {
  try { return static_cast<ref_model*>(h)->brightness_gpu(n, loc, dir, n_subsamples, out); } catch (const std::exception &e) { fprintf(stderr, "%s\n", e.what()); return -3; }
}
int ref_save_S(void *h, const char *fname) { return static_cast<ref_model*>(h)->save_S(fname); }
int ref_save_influence(void *h, const char *fname) { return static_cast<ref_model*>(h)->save_influence(fname); }
int ref_omp_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1: the caller pins the thread count of the CPU arm explicitly
void ref_set_omp_threads(int n) { if (n > 0) omp_set_num_threads(n); }

}
