// los_geom.cuh -- geometry of one line-of-sight sub-step, shared by the singlet (brightness.cu) and multiplet
// (multiplet.cu) brightness kernels.  Restates (reference src/):
//   atmo_vector::extend + atmo_point::xyz                  atmo_vec.cpp:292-306, 51-61
//   spherical_azimuthally_symmetric_grid::interp_weights   grid/grid_spherical_azimuthally_symmetric.hpp:511-612
#pragma once
#include "fastmath.cuh"

namespace b200rt {

template <class Real> struct MathB;
template <> struct MathB<double> {
  __device__ static double exp_(double x) { return fm::exp_nonpos(x); }   // arguments are <= 0 and finite
  __device__ static double div_(double a, double b) { return fm::div_approx(a, b); }   // b >= 1e-3 where it is used
  __device__ static double divq_(double a, double b) { return fm::div_fast(a, b); }    // 1e-12: inside the wavelength loops
  __device__ static double divc_(double a, double b, double rb) { return fm::div_by(a, b, rb); }   // rb = 1/b precomputed
  __device__ static double rcp_(double b) { return 1.0 / b; }
  __device__ static double log_(double x) { return log(x); }
  // the coordinates are O(1) after the 1e9 scaling: no overflow guards needed
  __device__ static double hypot2_(double a, double b, double c) { return fm::norm3(a, b, c); }
  __device__ static double acos_(double x) { return acos(x); }
  __device__ static double eps() { return 1e-6; }       // EPS      Real.hpp:23
  __device__ static double coneeps() { return 1e-6; }   // CONEEPS  Real.hpp:25
};
template <> struct MathB<float> {
  __device__ static float exp_(float x) { return expf(x); }
  __device__ static float div_(float a, float b) { return a / b; }
  __device__ static float divq_(float a, float b) { return a / b; }
  __device__ static float divc_(float a, float b, float) { return a / b; }
  __device__ static float rcp_(float b) { return 1.0f / b; }
  // std::log(float) of the host libm is (nearly) correctly rounded; CUDA logf is not (1 ulp), and one ulp of
  // logf(r) moves the radial interpolation weight by ~3e-5.  Rounding the double log gives the host's result.
  __device__ static float log_(float x) { return (float) log((double) x); }
  // atmo_point::xyz calls the unqualified (double) hypot / acos even when Real = float
  // (atmo_vec.cpp:53-54) and rounds on assignment: do the same
  __device__ static float hypot2_(float a, float b, float c) { return (float) hypot(hypot((double) a, (double) b), (double) c); }
  __device__ static float acos_(float x) { return (float) acos((double) x); }
  __device__ static float eps() { return 1e-3f; }       // Real.hpp:14
  __device__ static float coneeps() { return 1e-2f; }   // Real.hpp:16
};


// point at distance `dist` along the line of sight (position already divided by the 1e9 scale) -> the four
// neighbour voxels and bilinear weights (log r x linear SZA) of the interpolation inside voxel `cur`
template <class Real>
__device__ __forceinline__ void substep_interp(const Real *__restrict__ s_rb, const Real *__restrict__ s_sb,
                                               const Real *__restrict__ s_pr, const Real *__restrict__ s_lpr,
                                               const Real *__restrict__ s_ps, int n_rb, int n_sb1, int cur, Real px,
                                               Real py, Real pz, Real lx, Real ly, Real lz, Real dist, int (&idx)[4],
                                               Real (&w)[4]) {
  const Real eps = MathB<Real>::eps(), ceps = MathB<Real>::coneeps();
  const Real scale = Real(1e9);
  const Real r_scale = MathB<Real>::rcp_(scale);
  const int r_idx = cur / n_sb1, sza_idx = cur - r_idx * n_sb1;
  const Real rb_lo = s_rb[r_idx], rb_hi = s_rb[r_idx + 1];
  const Real sb_lo = s_sb[sza_idx], sb_hi = s_sb[sza_idx + 1];
  // ---- atmo_vector::extend
  const Real nx = px + MathB<Real>::divc_(lx * dist, scale, r_scale);
  const Real ny = py + MathB<Real>::divc_(ly * dist, scale, r_scale);
  const Real nz = pz + MathB<Real>::divc_(lz * dist, scale, r_scale);
  const Real rr = MathB<Real>::hypot2_(nx, ny, nz);
  Real t = MathB<Real>::acos_(MathB<Real>::div_(nz, rr));
  Real r = rr * scale;
  // ---- interp_weights
  if (r < rb_lo && rb_lo / r > (1 - eps)) r = rb_lo + eps;
  if (rb_hi < r && r / rb_hi < (1 + eps)) r = rb_hi - eps;
  if (t < sb_lo && sb_lo / t > (1 - ceps)) t = sb_lo + ceps;
  if (sb_hi < t && t / sb_hi < (1 + ceps)) t = sb_hi - ceps;
  int rlo, rhi;
  Real r_wt;
  if (r_idx == 0 && r <= s_pr[0]) { rlo = rhi = 0; r_wt = 1.0; }
  else if (r_idx == n_rb - 2 && s_pr[n_rb - 2] <= r) { rlo = rhi = n_rb - 2; r_wt = 0.0; }
  else {
    rlo = (r < s_pr[r_idx]) ? r_idx - 1 : r_idx;
    rhi = rlo + 1;
    const Real l0 = s_lpr[rlo], l1 = s_lpr[rhi];
    r_wt = MathB<Real>::div_(MathB<Real>::log_(r) - l0, l1 - l0);
  }
  int slo = (t < s_ps[sza_idx]) ? sza_idx - 1 : sza_idx;
  slo = max(0, min(slo, n_sb1 - 2));          // guard (the reference would index out of bounds)
  const int shi = slo + 1;
  const Real p0 = s_ps[slo], p1 = s_ps[shi];
  const Real s_wt = MathB<Real>::div_(t - p0, p1 - p0);
  idx[0] = rlo * n_sb1 + slo; w[0] = (Real(1.0) - r_wt) * (Real(1.0) - s_wt);
  idx[1] = rhi * n_sb1 + slo; w[1] = r_wt * (Real(1.0) - s_wt);
  idx[2] = rlo * n_sb1 + shi; w[2] = (Real(1.0) - r_wt) * s_wt;
  idx[3] = rhi * n_sb1 + shi; w[3] = r_wt * s_wt;
}

} // namespace b200rt
