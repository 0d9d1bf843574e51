// los_geom.cuh -- geometry of one line-of-sight sub-step, shared by the singlet (brightness.cu) and multiplet
// (multiplet.cu) brightness kernels.  Restates (reference src/):
//   atmo_vector::extend + atmo_point::xyz                  atmo_vec.cpp:292-306, 51-61
//   spherical_azimuthally_symmetric_grid::interp_weights   grid/grid_spherical_azimuthally_symmetric.hpp:511-612
#pragma once
#include "fastmath.cuh"

namespace b200rt {

namespace fm {
__device__ __forceinline__ double rsqrt_pos(double x);
__device__ __forceinline__ double acos_fast(double x);
__device__ __forceinline__ double log_pos(double x);
}
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
  template <class TT> __device__ static double sza_weight(double t, const TT &T, int slo) { return (t - T.ps[slo]) * T.ips[slo]; }
  __device__ static double eps() { return 1e-6; }       // EPS      Real.hpp:23
  __device__ static double coneeps() { return 1e-6; }   // CONEEPS  Real.hpp:25
};
template <> struct MathB<float> {
  // inside the wavelength loops: ex2.approx / rcp.approx based (2 ulp + |x| 6e-8 on exp, 2 ulp on the quotient),
  // against the 1e-4 bar of the float build; everything that decides an index or a weight keeps IEEE operations
  __device__ static float exp_(float x) { return __expf(x); }
  __device__ static float div_(float a, float b) { return a / b; }
  __device__ static float divq_(float a, float b) { return __fdividef(a, b); }
  __device__ static float divc_(float a, float b, float) { return a / b; }
  __device__ static float rcp_(float b) { return 1.0f / b; }
  // std::log(float) of the host libm is (nearly) correctly rounded; CUDA logf is not (1 ulp), and one ulp of
  // logf(r) moves the radial interpolation weight by ~3e-5.  Rounding the double log gives the host's result.
  __device__ static float log_(float x) { return (float) fm::log_pos((double) x); }
  // atmo_point::xyz calls the unqualified (double) hypot / acos even when Real = float
  // (atmo_vec.cpp:53-54) and rounds on assignment: do the same
  // (the double intermediates are formed with the trimmed routines of the double path, good to <= 1e-13: after the
  // rounding to float they are the libm values in all but ~1e-6 of the cases, against a 1e-4 tolerance)
  __device__ static float hypot2_(float a, float b, float c) {
    const double r2 = fma((double) a, (double) a, fma((double) b, (double) b, (double) c * (double) c));
    return (float) (r2 * fm::rsqrt_pos(r2));
  }
  __device__ static float acos_(float x) { return (float) fm::acos_fast(fmin(fmax((double) x, -1.0), 1.0)); }
  template <class TT> __device__ static float sza_weight(float t, const TT &T, int slo) { return (t - T.ps[slo]) / (T.ps[slo + 1] - T.ps[slo]); }
  __device__ static float eps() { return 1e-3f; }       // Real.hpp:14
  __device__ static float coneeps() { return 1e-2f; }   // Real.hpp:16
};


// ---- the grid axes a block keeps in shared memory for the interpolation
template <class Real>
struct GeomTables {
  Real *rb, *sb, *pr, *lpr, *ps;      // boundaries, voxel points, log of the radial points
  Real *ipr, *ilw, *ips;              // 1/pr[i], 1/(lpr[i+1]-lpr[i]), 1/(ps[j+1]-ps[j]): the double path multiplies
  int n_rb, n_sb1;
  int narrow;                         // every radial cell ratio pr[i+1]/pr[i] <= 2: log(r/pr) by its atanh series
  static __host__ __device__ size_t doubles(int n_rb, int n_sb) { return (size_t) 6 * n_rb + 3 * n_sb; }
  // all threads of the block; ends with __syncthreads()
  __device__ void load(unsigned char *smem, const GridView<Real> &g) {
    n_rb = g.n_rb; n_sb1 = g.n_sb - 1;
    rb = reinterpret_cast<Real *>(smem); sb = rb + g.n_rb; pr = sb + g.n_sb; lpr = pr + g.n_rb; ps = lpr + g.n_rb;
    ipr = ps + g.n_sb; ilw = ipr + g.n_rb; ips = ilw + g.n_rb;
    for (int i = threadIdx.x; i < g.n_rb; i += blockDim.x) rb[i] = g.rb[i];
    for (int i = threadIdx.x; i < g.n_sb; i += blockDim.x) sb[i] = g.sb[i];
    for (int i = threadIdx.x; i < g.n_rb - 1; i += blockDim.x) {
      pr[i] = g.pts_r[i]; lpr[i] = g.log_pts_r[i];
      ipr[i] = Real(1) / g.pts_r[i];
      ilw[i] = (i < g.n_rb - 2) ? Real(1) / (g.log_pts_r[i + 1] - g.log_pts_r[i]) : Real(0);
    }
    for (int i = threadIdx.x; i < n_sb1; i += blockDim.x) {
      ps[i] = g.pts_s[i];
      ips[i] = (i < n_sb1 - 1) ? Real(1) / (g.pts_s[i + 1] - g.pts_s[i]) : Real(0);
    }
    __syncthreads();
    bool ok = true;
    for (int i = 0; i < g.n_rb - 2; i++) ok = ok && (pr[i + 1] <= Real(2) * pr[i]);
    narrow = ok ? 1 : 0;
  }
};

namespace fm {
// 1/sqrt(x) for normal positive x, ~1e-15 relative
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  return y;
}
// acos(x), |x| <= 1, absolute error < 1e-13: acos(a) = sqrt(1 - a) * P14(a) on [0, 1] (least-squares fit on Chebyshev
// nodes, max error 4.4e-14), reflected for x < 0.  The interpolation weight it feeds is held to 1e-6.
__constant__ double ACOSC[15] = {
    1.5707963267948528, -0.21460183658219414, 0.0890486209836325, -0.05079276681296366, 0.033680519262737484,
    -0.02436683099115829, 0.01862394790520326, -0.014669083847799712, 0.011512617319082787, -0.008535619736430394,
    0.005561810086541537, -0.0029281197388131013, 0.0011292831396020055, -0.000277475668365072, 3.217025920001402e-05};
__device__ __forceinline__ double acos_fast(double x) {
  const double a = fabs(x);
  const double om = fmax(1.0 - a, 0.0);
  const double s = (om > 1e-300) ? om * rsqrt_pos(om) : 0.0;
  double p = ACOSC[14];
#pragma unroll
  for (int i = 13; i >= 0; i--) p = fma(p, a, ACOSC[i]);
  const double r = s * p;
  return (x >= 0.0) ? r : 3.14159265358979323846 - r;
}
// log(1 + u) for -0.05 <= u <= 1 (z = u / (2 + u) <= 1/3): 2 atanh(z) by its series to z^15, error < 1e-9 at u = 1 and
// < 1e-13 for u <= 0.5 (cell ratios of the grids in use are <= 1.32)
__device__ __forceinline__ double log1p_series(double u) {
  const double z = div_approx(u, 2.0 + u);
  const double z2 = z * z;
  double p = 1.0 / 15.0;
  p = fma(p, z2, 1.0 / 13.0);
  p = fma(p, z2, 1.0 / 11.0);
  p = fma(p, z2, 1.0 / 9.0);
  p = fma(p, z2, 1.0 / 7.0);
  p = fma(p, z2, 1.0 / 5.0);
  p = fma(p, z2, 1.0 / 3.0);
  p = fma(p, z2, 1.0);
  return 2.0 * z * p;
}
// log(x) for normal positive x: exponent split + the series on the mantissa in [0.75, 1.5), error < 1e-12
__device__ __forceinline__ double log_pos(double x) {
  int hi = __double2hiint(x);
  int e = ((hi >> 20) & 0x7ff) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
  if (m > 1.5) { m *= 0.5; e += 1; }
  return fma((double) e, 0.69314718055994530942, log1p_series(m - 1.0));
}
} // namespace fm

// point at distance `dist` along the line of sight (position already divided by the 1e9 scale) -> the four
// neighbour voxels and bilinear weights (log r x linear SZA) of the interpolation inside voxel `cur`.
// Common tail: indices and weights from (r, t, log-r weight source).
template <class Real, class LogW>
__device__ __forceinline__ void interp_from_rt(const GeomTables<Real> &T, int cur, Real r, Real t, LogW log_weight,
                                               int (&idx)[4], Real (&w)[4]) {
  const Real eps = MathB<Real>::eps(), ceps = MathB<Real>::coneeps();
  const int n_rb = T.n_rb, n_sb1 = T.n_sb1;
  const int r_idx = cur / n_sb1, sza_idx = cur - r_idx * n_sb1;
  const Real rb_lo = T.rb[r_idx], rb_hi = T.rb[r_idx + 1];
  const Real sb_lo = T.sb[sza_idx], sb_hi = T.sb[sza_idx + 1];
  // ---- interp_weights
  if (r < rb_lo && rb_lo / r > (1 - eps)) r = rb_lo + eps;
  if (rb_hi < r && r / rb_hi < (1 + eps)) r = rb_hi - eps;
  if (t < sb_lo && sb_lo / t > (1 - ceps)) t = sb_lo + ceps;
  if (sb_hi < t && t / sb_hi < (1 + ceps)) t = sb_hi - ceps;
  int rlo, rhi;
  Real r_wt;
  if (r_idx == 0 && r <= T.pr[0]) { rlo = rhi = 0; r_wt = 1.0; }
  else if (r_idx == n_rb - 2 && T.pr[n_rb - 2] <= r) { rlo = rhi = n_rb - 2; r_wt = 0.0; }
  else {
    rlo = (r < T.pr[r_idx]) ? r_idx - 1 : r_idx;
    rhi = rlo + 1;
    r_wt = log_weight(r, rlo);
  }
  int slo = (t < T.ps[sza_idx]) ? sza_idx - 1 : sza_idx;
  slo = max(0, min(slo, n_sb1 - 2));          // guard (the reference would index out of bounds)
  const int shi = slo + 1;
  const Real s_wt = MathB<Real>::sza_weight(t, T, slo);
  idx[0] = rlo * n_sb1 + slo; w[0] = (Real(1.0) - r_wt) * (Real(1.0) - s_wt);
  idx[1] = rhi * n_sb1 + slo; w[1] = r_wt * (Real(1.0) - s_wt);
  idx[2] = rlo * n_sb1 + shi; w[2] = (Real(1.0) - r_wt) * s_wt;
  idx[3] = rhi * n_sb1 + shi; w[3] = r_wt * s_wt;
}

template <class Real>
__device__ __forceinline__ void substep_interp(const GeomTables<Real> &T, int cur, Real px, Real py, Real pz, Real lx,
                                               Real ly, Real lz, Real dist, int (&idx)[4], Real (&w)[4]);

// float: the reference's operations one for one (hypot / acos / log in double, rounded), so that the float build lands
// on the host's values (util.TOL_AUX)
template <>
__device__ __forceinline__ void substep_interp<float>(const GeomTables<float> &T, int cur, float px, float py, float pz,
                                                      float lx, float ly, float lz, float dist, int (&idx)[4], float (&w)[4]) {
  typedef float Real;
  const Real scale = Real(1e9);
  const Real r_scale = MathB<Real>::rcp_(scale);
  // ---- atmo_vector::extend
  const Real nx = px + MathB<Real>::divc_(lx * dist, scale, r_scale);
  const Real ny = py + MathB<Real>::divc_(ly * dist, scale, r_scale);
  const Real nz = pz + MathB<Real>::divc_(lz * dist, scale, r_scale);
  const Real rr = MathB<Real>::hypot2_(nx, ny, nz);
  const Real t = MathB<Real>::acos_(MathB<Real>::div_(nz, rr));
  const Real r = rr * scale;
  interp_from_rt<Real>(T, cur, r, t, [&](Real rv, int rlo) {
    const Real l0 = T.lpr[rlo], l1 = T.lpr[rlo + 1];
    return MathB<Real>::div_(MathB<Real>::log_(rv) - l0, l1 - l0);
  }, idx, w);
}

// double: the same geometry with the transcendental work trimmed to what a 1e-6 tolerance on the brightness needs
// (every piece below is good to <= 1e-13): 1/|p| by rsqrt + two Newton steps (no sqrt, no division), acos by
// sqrt(1-a) P14(a), the radial weight log(r / pr[rlo]) / log(pr[rlo+1] / pr[rlo]) by the atanh series of log(1+u) with
// tabulated reciprocals.  ~65 of the ~180 FP64 instructions of a sub-step's geometry go away (r01k: 51.6 -> see profiles).
template <>
__device__ __forceinline__ void substep_interp<double>(const GeomTables<double> &T, int cur, double px, double py, double pz,
                                                       double lx, double ly, double lz, double dist, int (&idx)[4],
                                                       double (&w)[4]) {
  const double ds = dist * 1e-9;
  const double nx = fma(lx, ds, px), ny = fma(ly, ds, py), nz = fma(lz, ds, pz);
  const double r2 = fma(nx, nx, fma(ny, ny, nz * nz));
  const double inv = fm::rsqrt_pos(r2);
  const double t = fm::acos_fast(fmin(fmax(nz * inv, -1.0), 1.0));
  const double r = (r2 * inv) * 1e9;
  if (T.narrow)
    interp_from_rt<double>(T, cur, r, t, [&](double rv, int rlo) {
      return fm::log1p_series(fma(rv, T.ipr[rlo], -1.0)) * T.ilw[rlo];
    }, idx, w);
  else
    interp_from_rt<double>(T, cur, r, t, [&](double rv, int rlo) {
      return (log(rv) - T.lpr[rlo]) * T.ilw[rlo];
    }, idx, w);
}

} // namespace b200rt
