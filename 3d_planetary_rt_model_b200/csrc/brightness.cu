// brightness.cu -- line-of-sight brightness integration (sm_100a).
//
// Restates, for the device (reference src/):
//   RT_grid::brightness(vec, los, n_subsamples)            RT_grid.hpp:233-299
//   atmo_vector::extend + atmo_point::xyz                  atmo_vec.cpp:292-306, 51-61
//   spherical_azimuthally_symmetric_grid::interp_weights   grid/grid_spherical_azimuthally_symmetric.hpp:511-612
//   emission_voxels::update_tracker_brightness_{interp,nointerp}, interp_voxel_vector
//                                                           emission/emission_voxels.hpp:58-70,199-233
//   singlet_CFR::update_tracker_start<false>, update_tracker_start_interp, update_tracker_brightness
//                                                           emission/singlet_CFR.hpp:80-260,315-348,262-276
//   los_tracker::exits_bottom                               emission/los_tracker.hpp:57-62
//
// The boundary list of each line of sight comes from traverse.cu.  Mapping: one
// thread per line of sight (the reference kernel does the same with 32-thread
// blocks and a 232-byte tracker staged through shared memory; here the tracker is
// the register file: P[20] per emission, fully unrolled), all emissions marched in
// the same pass so the geometry (extend, interpolation weights) is done once.
// Inputs and outputs are SoA so every global access of a warp is coalesced.
#include "common.hpp"

namespace b200rt {

namespace {

template <class Real> struct MathB;
template <> struct MathB<double> {
  __device__ static double exp_(double x) { return exp(x); }
  __device__ static double log_(double x) { return log(x); }
  __device__ static double hypot2_(double a, double b, double c) { return hypot(hypot(a, b), c); }
  __device__ static double acos_(double x) { return acos(x); }
  __device__ static double eps() { return 1e-6; }       // EPS      Real.hpp:23
  __device__ static double coneeps() { return 1e-6; }   // CONEEPS  Real.hpp:25
};
template <> struct MathB<float> {
  __device__ static float exp_(float x) { return expf(x); }
  __device__ static float log_(float x) { return logf(x); }
  // atmo_point::xyz calls the unqualified (double) hypot / acos even when Real = float
  // (atmo_vec.cpp:53-54) and rounds on assignment: do the same
  __device__ static float hypot2_(float a, float b, float c) { return (float) hypot(hypot((double) a, (double) b), (double) c); }
  __device__ static float acos_(float x) { return (float) acos((double) x); }
  __device__ static float eps() { return 1e-3f; }       // Real.hpp:14
  __device__ static float coneeps() { return 1e-2f; }   // Real.hpp:16
};

template <class Real>
struct Tracker {   // brightness_tracker + singlet_CFR_tracker<false> (los_tracker.hpp:11-170)
  Real tau_sp, tau_abs, col, B;
  Real P[N_LAMBDA];
};

// singlet_CFR::update_tracker_start<false> + update_tracker_brightness
template <class Real>
__device__ __forceinline__ void step(Tracker<Real> &t, Real Tr, Real dens, Real dts, Real dta, Real s, Real Sv,
                                     Real gfac) {
  t.col += dens * s;
  const Real tau_species_voxel = dts * s;
  t.tau_sp += tau_species_voxel;
  t.tau_abs += dta * s;
  const Real delta_lambda = Real(4.0) / (N_LAMBDA - 1);
  const Real common = dts * s;
  Real T_int = 0;
#pragma unroll
  for (int i = 0; i < N_LAMBDA; i++) {
    const Real l = i * delta_lambda;
    const Real lineshape = MathB<Real>::exp_(-(l * l) * Tr);
    const Real tau = (dta + dts * lineshape) * s;
    const Real tp = MathB<Real>::exp_(-tau);
    const Real wgt = (i == 0 || i == N_LAMBDA - 1) ? delta_lambda : Real(2.0) * delta_lambda;
    Real c = ((double) tau < 1e-3) ? (Real(1.0) - Real(0.5) * tau) : (Real(1.0) - tp) / tau;
    c *= (wgt * lineshape * t.P[i]) * common;
    T_int += c;
    t.P[i] *= tp;
  }
  if (T_int > tau_species_voxel) T_int = tau_species_voxel;
  t.B += Sv * gfac * T_int;     // gfac = g * branching / sigma_ref / sqrt(pi) / 1e9
}

template <class Real>
__device__ __forceinline__ Real interp4(const Real *__restrict__ q, const int (&idx)[4], const Real (&w)[4]) {
  Real s = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += w[k] * q[idx[k]];
  return s;
}

template <class Real, int NEM>
__global__ void __launch_bounds__(128)
brightness_kernel(GridView<Real> g, EmissionView<Real> em0, EmissionView<Real> em1,
                  const Real *__restrict__ los_in, long long los_stride, long long first, long long count,
                  ListView<Real> lists, int n_subsamples, Real *__restrict__ out, long long n_los_total) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const long long los = first + i;
  const EmissionView<Real> em[2] = {em0, em1};

  const Real px = los_in[0 * los_stride + los], py = los_in[1 * los_stride + los], pz = los_in[2 * los_stride + los];
  const Real lx = los_in[5 * los_stride + los], ly = los_in[6 * los_stride + los], lz = los_in[7 * los_stride + los];

  Tracker<Real> tr[NEM];
  Real gfac[NEM];
#pragma unroll
  for (int e = 0; e < NEM; e++) {
    tr[e].tau_sp = 0; tr[e].tau_abs = 0; tr[e].col = 0; tr[e].B = 0;
#pragma unroll
    for (int k = 0; k < N_LAMBDA; k++) tr[e].P[k] = Real(1);
    gfac[e] = em[e].g_factor * em[e].branching / em[e].sigma_ref * (Real) 0.56418958354775628695 / Real(1e9);
  }

  const int len = lists.len[i];
  const int n_sb1 = g.n_sb - 1;
  if (len > 0) {
    const Real *dl = lists.dist + (size_t) i * lists.cap;
    const int *el = lists.ent + (size_t) i * lists.cap;
    const int nsd = (n_subsamples == 0) ? 2 : n_subsamples;
    const Real eps = MathB<Real>::eps(), ceps = MathB<Real>::coneeps();
    const Real scale = Real(1e9);
    Real dprev = dl[0];
    int cur = el[0];
    for (int ib = 1; ib < len; ib++) {
      const Real dnext = dl[ib];
      Real d_start = dprev;
      Real d_step = (dnext - d_start) / (nsd - 1);
      d_start += Real(0.5) * eps * d_step;          // RT_grid.hpp:268-271
      d_step *= Real(1.0) - eps;

      if (n_subsamples == 0) {
#pragma unroll
        for (int e = 0; e < NEM; e++)
          step(tr[e], em[e].T_ratio[cur], em[e].density[cur], em[e].dtau_species[cur], em[e].dtau_absorber[cur],
               d_step, em[e].sourcefn[cur], gfac[e]);
      } else {
        const int r_idx = cur / n_sb1, sza_idx = cur % n_sb1;
        const Real rb_lo = g.rb[r_idx], rb_hi = g.rb[r_idx + 1];
        const Real sb_lo = g.sb[sza_idx], sb_hi = g.sb[sza_idx + 1];
        const Real ps_c = g.pts_s[sza_idx];
        for (int is = 1; is < nsd; is++) {
          // ---- atmo_vector::extend
          const Real dist = d_start + is * d_step;
          const Real nx = px / scale + (lx * dist) / scale;
          const Real ny = py / scale + (ly * dist) / scale;
          const Real nz = pz / scale + (lz * dist) / scale;
          const Real rr = MathB<Real>::hypot2_(nx, ny, nz);
          Real t = MathB<Real>::acos_(nz / rr);
          Real r = rr * scale;
          // ---- interp_weights
          if (r < rb_lo && rb_lo / r > (1 - eps)) r = rb_lo + eps;
          if (rb_hi < r && r / rb_hi < (1 + eps)) r = rb_hi - eps;
          if (t < sb_lo && sb_lo / t > (1 - ceps)) t = sb_lo + ceps;
          if (sb_hi < t && t / sb_hi < (1 + ceps)) t = sb_hi - ceps;
          int rlo, rhi;
          Real r_wt;
          if (r_idx == 0 && r <= g.pts_r[0]) { rlo = rhi = 0; r_wt = 1.0; }
          else if (r_idx == g.n_rb - 2 && g.pts_r[g.n_rb - 2] <= r) { rlo = rhi = g.n_rb - 2; r_wt = 0.0; }
          else {
            rlo = (r < g.pts_r[r_idx]) ? r_idx - 1 : r_idx;
            rhi = rlo + 1;
            const Real l0 = g.log_pts_r[rlo], l1 = g.log_pts_r[rhi];
            r_wt = (MathB<Real>::log_(r) - l0) / (l1 - l0);
          }
          int slo = (t < ps_c) ? sza_idx - 1 : sza_idx;
          slo = max(0, min(slo, n_sb1 - 2));        // guard (the reference would index out of bounds)
          const int shi = slo + 1;
          const Real p0 = g.pts_s[slo], p1 = g.pts_s[shi];
          const Real s_wt = (t - p0) / (p1 - p0);
          int idx[4];
          Real w[4];
          idx[0] = rlo * n_sb1 + slo; w[0] = (Real(1.0) - r_wt) * (Real(1.0) - s_wt);
          idx[1] = rhi * n_sb1 + slo; w[1] = r_wt * (Real(1.0) - s_wt);
          idx[2] = rlo * n_sb1 + shi; w[2] = (Real(1.0) - r_wt) * s_wt;
          idx[3] = rhi * n_sb1 + shi; w[3] = r_wt * s_wt;
#pragma unroll
          for (int e = 0; e < NEM; e++) {
            const Real Tr = interp4(em[e].T_ratio_pt, idx, w);
            const Real dn = interp4(em[e].density_pt, idx, w);
            const Real ds = interp4(em[e].dtau_species_pt, idx, w);
            const Real da = interp4(em[e].dtau_absorber_pt, idx, w);
            const Real Sv = interp4(em[e].sourcefn, idx, w);
            step(tr[e], Tr, dn, ds, da, d_step, Sv, gfac[e]);
          }
        }
      }
      dprev = dnext;
      cur = el[ib];
    }
    if (lists.flag[i] & 1) {
#pragma unroll
      for (int e = 0; e < NEM; e++) tr[e].tau_abs = Real(-1.0);   // los_tracker::exits_bottom
    }
  }
#pragma unroll
  for (int e = 0; e < NEM; e++) {
    out[((size_t) e * 4 + 0) * n_los_total + los] = tr[e].B;
    out[((size_t) e * 4 + 1) * n_los_total + los] = tr[e].tau_sp;
    out[((size_t) e * 4 + 2) * n_los_total + los] = tr[e].tau_abs;
    out[((size_t) e * 4 + 3) * n_los_total + los] = tr[e].col;
  }
}

} // namespace

template <class Real>
cudaError_t launch_brightness(const GridView<Real> &g, const EmissionView<Real> *em, int n_em, const Real *los_in,
                              long long los_stride, long long first, long long count, ListView<Real> lists,
                              int n_subsamples, Real *out, long long n_los_total, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const int threads = 128;
  const unsigned blocks = (unsigned) ((count + threads - 1) / threads);
  if (n_em == 1)
    brightness_kernel<Real, 1><<<blocks, threads, 0, s>>>(g, em[0], em[0], los_in, los_stride, first, count, lists,
                                                          n_subsamples, out, n_los_total);
  else
    brightness_kernel<Real, 2><<<blocks, threads, 0, s>>>(g, em[0], em[1], los_in, los_stride, first, count, lists,
                                                          n_subsamples, out, n_los_total);
  return cudaGetLastError();
}
template cudaError_t launch_brightness<double>(const GridView<double> &, const EmissionView<double> *, int,
                                               const double *, long long, long long, long long, ListView<double>, int,
                                               double *, long long, cudaStream_t);
template cudaError_t launch_brightness<float>(const GridView<float> &, const EmissionView<float> *, int, const float *,
                                              long long, long long, long long, ListView<float>, int, float *,
                                              long long, cudaStream_t);

} // namespace b200rt
