// RT_b200.hpp -- the reference-side binding of libb200rt.so: definitions of the RT_grid members the reference DECLARES
// in RT_grid.hpp:31-39,146,219,325 and defines only in src/RT_gpu.cu (:8-84,138-192,255-309), written as plain C++
// that marshals to the C ABI of include/b200rt.h.  A maintainer includes this header after RT_grid.hpp in the one
// translation unit that calls generate_S_gpu() / brightness_gpu() (observation_fit.cpp), compiles with the host
// compiler (no nvcc, RT_gpu.cu is no longer pulled in) and links -lb200rt.  Nothing in the reference's headers changes;
// d_RT / d_obs / device_emission simply stay NULL.
//
// This file is compiled and exercised by oracle/Makefile (target ref_b200) against the reference's own headers in
// /root/reference/src: tests/test_rt_grid_stub.py calls RT_grid::generate_S_gpu() and brightness_gpu() through it and
// compares with the same object's CPU generate_S() / brightness().
//
// Scope: RT_grid<singlet_CFR<N_VOXELS>, N_EMISSIONS, spherical_azimuthally_symmetric_grid<...>> -- the type
// observation_fit instantiates as hydrogen_RT / deuterium_RT (observation_fit.hpp:56-76).
#ifndef RT_B200_HPP
#define RT_B200_HPP

#include <stdexcept>
#include <string>
#include <vector>
#include "RT_grid.hpp"
#include "emission/singlet_CFR.hpp"
#include "b200rt.h"

namespace rt_b200 {

// one context per host thread, created on first use (every visible GPU behind one handle), destroyed with the thread
struct ctx_holder {
  b200rt_ctx *c = nullptr;
  ~ctx_holder() { if (c) b200rt_destroy(c); }
  b200rt_ctx *get() {
    if (!c && b200rt_create_multi(0, nullptr, sizeof(Real) == 8 ? B200RT_F64 : B200RT_F32, &c) != B200RT_OK)
      throw std::runtime_error("b200rt: no usable CUDA device (there is no CPU fallback)");
    return c;
  }
};
// (keyed by Real: function-local statics of inline functions are process-wide unique symbols, so a process that loads
// a float and a double build of the reference side by side must not share one context between them)
template <class R> inline ctx_holder &holder_of() { static thread_local ctx_holder h; return h; }
inline ctx_holder &holder() { return holder_of<Real>(); }
inline void ck(int rc) {
  if (rc != B200RT_OK) throw std::runtime_error(std::string("b200rt status ") + std::to_string(rc) + ": " + b200rt_last_error(holder().c));
}

// singlet_CFR keeps its per-voxel tables and the emission base its flags protected.  A pointer to an inherited member
// named through a derived class has the BASE's pointer-to-member type, so it applies to the reference's own objects:
// no friend declaration and no change to the reference's headers is needed.
template <int NV>
struct peek : singlet_CFR<NV> {
  typedef singlet_CFR<NV> E;
  static std::vector<double> table(const E &e, typename E::vv E::*m) {
    std::vector<double> out(NV);
    for (int i = 0; i < NV; i++) out[i] = (double) (e.*m)(i);
    return out;
  }
#define RT_B200_TABLE(name) static std::vector<double> name##_of(const E &e) { return table(e, &peek::name); }
  RT_B200_TABLE(species_T_ratio) RT_B200_TABLE(species_density) RT_B200_TABLE(dtau_species) RT_B200_TABLE(dtau_absorber)
  RT_B200_TABLE(species_T_ratio_pt) RT_B200_TABLE(species_density_pt) RT_B200_TABLE(dtau_species_pt) RT_B200_TABLE(dtau_absorber_pt)
#undef RT_B200_TABLE
  static double branching_ratio_of(const E &e) { return e.*(&peek::branching_ratio); }
  static double T_ref_of(const E &e) { return e.*(&peek::species_T_ref); }
  static double sigma_T_ref_of(const E &e) { return e.*(&peek::species_sigma_T_ref); }
  static double g_factor_of(const E &e) { return e.*(&peek::emission_g_factor); }
  static std::vector<double> sourcefn_of(const E &e) {
    std::vector<double> out(NV);
    for (int i = 0; i < NV; i++) out[i] = (double) (e.*(&peek::sourcefn))(i);
    return out;
  }
  static void store_solution(E &e, const double *S, const double *S0, const double *ts, const double *ta) {
    for (int i = 0; i < NV; i++) {
      (e.*(&peek::sourcefn))(i) = (Real) S[i];
      (e.*(&peek::singlescat))(i) = (Real) S0[i];
      (e.*(&peek::tau_species_single_scattering))(i) = (Real) ts[i];
      (e.*(&peek::tau_absorber_single_scattering))(i) = (Real) ta[i];
    }
    e.*(&peek::internal_solved) = true;      // emission_voxels::solve, emission_voxels.hpp:194
  }
  static void store_influence(E &e, const double *K) {   // row-major [NV][NV] (b200rt_get_influence)
    for (int i = 0; i < NV; i++)
      for (int j = 0; j < NV; j++) (e.*(&peek::influence_matrix))(i, j) = (Real) K[(size_t) i * NV + j];
  }
};

template <class T>
inline std::vector<double> widen(const T *p, int n) { return std::vector<double>(p, p + n); }

}  // namespace rt_b200

// replaces RT_grid::RT_to_device + emissions_to_device_* (RT_gpu.cu:8-60; singlet_CFR::copy_to_device_*,
// singlet_CFR.hpp:565-613): geometry and the eight per-voxel tables of every emission
template <typename E, int N, typename G>
void RT_grid<E, N, G>::RT_to_device() {
  using namespace rt_b200;
  typedef peek<G::n_voxels> P;
  b200rt_ctx *c = holder().get();
  const G &g = grid;     // members of spherical_azimuthally_symmetric_grid, grid_spherical_azimuthally_symmetric.hpp:47-73
  std::vector<double> rt(G::n_rays), rp(G::n_rays), rw(G::n_rays);
  for (int i = 0; i < G::n_rays; i++) { rt[i] = g.rays[i].t; rp[i] = g.rays[i].p; rw[i] = g.rays[i].domega; }
  ck(b200rt_set_grid_sph(c, G::n_radial_boundaries, G::n_sza_boundaries, G::n_rays,
                         widen(g.radial_boundaries, G::n_radial_boundaries).data(),
                         widen(g.sza_boundaries, G::n_sza_boundaries).data(),
                         widen(g.pts_radii, G::n_radial_boundaries - 1).data(),
                         widen(g.pts_sza, G::n_sza_boundaries - 1).data(), rt.data(), rp.data(), rw.data()));
  for (int e = 0; e < N; e++) {
    const E &m = *emissions[e];
    ck(b200rt_set_singlet(c, e, N, P::branching_ratio_of(m), P::T_ref_of(m), P::sigma_T_ref_of(m), P::g_factor_of(m),
                          P::species_T_ratio_of(m).data(), P::species_density_of(m).data(), P::dtau_species_of(m).data(),
                          P::dtau_absorber_of(m).data(), P::species_T_ratio_pt_of(m).data(),
                          P::species_density_pt_of(m).data(), P::dtau_species_pt_of(m).data(),
                          P::dtau_absorber_pt_of(m).data()));
  }
}
template <typename E, int N, typename G> void RT_grid<E, N, G>::RT_to_device_influence() { RT_to_device(); }
template <typename E, int N, typename G> void RT_grid<E, N, G>::RT_to_device_brightness() {
  // geometry, tables and the source function of THIS object (the per-thread context may last have served another one)
  using namespace rt_b200;
  RT_to_device();
  for (int e = 0; e < N; e++) {
    ck(b200rt_set_sourcefn(holder().c, e, peek<G::n_voxels>::sourcefn_of(*emissions[e]).data()));
  }
}

// replaces RT_grid::generate_S_gpu (RT_gpu.cu:255-309): influence kernel + solve_gpu + emissions_solved_to_host
template <typename E, int N, typename G>
void RT_grid<E, N, G>::generate_S_gpu() {
  using namespace rt_b200;
  typedef peek<G::n_voxels> P;
  RT_to_device();
  b200rt_ctx *c = holder().c;
  ck(b200rt_generate_S(c));
  const int nv = G::n_voxels;
  std::vector<double> S(nv), S0(nv), ts(nv), ta(nv);
  for (int e = 0; e < N; e++) {
    ck(b200rt_get_solution(c, e, S.data(), S0.data(), ts.data(), ta.data()));
    P::store_solution(*emissions[e], S.data(), S0.data(), ts.data(), ta.data());
  }
}

// replaces RT_grid::emissions_influence_to_host (RT_gpu.cu:62-72): K stays resident on the device after generate_S_gpu
// and comes back only when the caller wants it (save_influence)
template <typename E, int N, typename G>
void RT_grid<E, N, G>::emissions_influence_to_host() {
  using namespace rt_b200;
  typedef peek<G::n_voxels> P;
  std::vector<double> K((size_t) G::n_voxels * G::n_voxels);
  for (int e = 0; e < N; e++) {
    ck(b200rt_get_influence(holder().get(), e, B200RT_ROW_MAJOR, K.data()));
    // the reference's influence_matrix holds the branching-ratio-scaled kernel after solve (pre_solve scales in place)
    const double w = P::branching_ratio_of(*emissions[e]);
    for (double &v : K) v *= w;
    P::store_influence(*emissions[e], K.data());
  }
}

// replaces RT_grid::brightness_gpu (RT_gpu.cu:138-192) and observation::to_device / to_host (observation.hpp:211-254)
template <typename E, int N, typename G>
void RT_grid<E, N, G>::brightness_gpu(observation<E, N> &obs, const int n_subsamples) {
  using namespace rt_b200;
  RT_to_device_brightness();   // as the reference does on every call (RT_gpu.cu:146-148): this object's tables and S
  const int n = obs.size();
  std::vector<double> a[9];
  for (auto &v : a) v.resize(n);
  for (int i = 0; i < n; i++) {                           // the fields of obs_vecs[i] (atmo_vec.hpp)
    const atmo_vector q = obs.get_vec(i);
    a[0][i] = q.pt.x; a[1][i] = q.pt.y; a[2][i] = q.pt.z; a[3][i] = q.pt.r; a[4][i] = q.pt.t;
    a[5][i] = q.line_x; a[6][i] = q.line_y; a[7][i] = q.line_z; a[8][i] = q.ray.cost;
  }
  std::vector<double> B((size_t) N * n), ts((size_t) N * n), ta((size_t) N * n), col((size_t) N * n);
  ck(b200rt_brightness(holder().c, n, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(),
                       a[6].data(), a[7].data(), a[8].data(), n_subsamples, B.data(), ts.data(), ta.data(), col.data()));
  for (int e = 0; e < N; e++)
    for (int i = 0; i < n; i++) {                         // the four tracker members observation_fit reads (:491-559)
      auto &t = obs.los[e][i];
      t.brightness = (Real) B[(size_t) e * n + i];
      t.tau_species_final = (Real) ts[(size_t) e * n + i];
      t.tau_absorber_final = (Real) ta[(size_t) e * n + i];
      t.species_col_dens = (Real) col[(size_t) e * n + i];
    }
}
template <typename E, int N, typename G> void RT_grid<E, N, G>::device_clear() {}   // the context owns all device memory

#endif
