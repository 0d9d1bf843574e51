// influence.cu -- Holstein influence-matrix march and single scattering (sm_100a).
//
// Restates, for the device (reference src/):
//   singlet_CFR::update_tracker_start<true>        emission/singlet_CFR.hpp:80-260
//   singlet_CFR::update_tracker_influence          emission/singlet_CFR.hpp:352-370
//   singlet_CFR::compute_single_scattering         emission/singlet_CFR.hpp:372-398
//   singlet_CFR_tracker (lambda grid, weights)     emission/los_tracker.hpp:117-170
//   RT_grid::generate_S loop body                  RT_grid.hpp:166-201
//   emission_voxels::accumulate_influence          emission/emission_voxels.hpp:137-168
//
// It consumes the boundary lists written by traverse.cu, so no geometry is decided
// here; FMA contraction is allowed (results differ from the reference by rounding,
// ~1e-16 relative, against a 1e-6 tolerance).
//
// Mapping: a warp marches 8 rays at once, 4 lanes per ray, each lane carrying 5 of the
// 20 wavelength points (lambda index = sub + 4 m): the per-wavelength transmission
// vector P[] lives in registers, the per-step sum over wavelengths is two shuffles,
// and the step's contribution domega*G goes to K[row][voxel] with one fp64 RED.
// Per-voxel line shapes phi(lambda_i; T) are tabulated once per emission
// (phi_table_kernel) so a step costs one exp() per wavelength instead of three.
#include "common.hpp"

namespace b200rt {

namespace {

template <class Real> __device__ __forceinline__ Real rexp(Real x);
template <> __device__ __forceinline__ double rexp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float rexp<float>(float x) { return expf(x); }
template <class Real> __device__ __forceinline__ Real rsqrt_(Real x);
template <> __device__ __forceinline__ double rsqrt_<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float rsqrt_<float>(float x) { return sqrtf(x); }

template <class Real>
__device__ __forceinline__ Real lambda2_of(int i) {
  // lambda(i) = i*delta_lambda, delta_lambda = lambda_max/(n_lambda-1)   (los_tracker.hpp:123-130)
  const Real delta_lambda = Real(4.0) / (N_LAMBDA - 1);
  Real l = i * delta_lambda;
  return l * l;
}
template <class Real>
__device__ __forceinline__ Real weight_of(int i) {
  const Real delta_lambda = Real(4.0) / (N_LAMBDA - 1);
  return (i == 0 || i == N_LAMBDA - 1) ? delta_lambda : Real(2.0) * delta_lambda;
}
template <class Real>
__device__ __forceinline__ Real one_over_sqrt_pi() { return (Real) 0.56418958354775628695; }

template <class Real>
__global__ void phi_table_kernel(const Real *__restrict__ T_ratio, int n_vox, Real *__restrict__ phi) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_vox * N_LAMBDA) return;
  const int v = idx / N_LAMBDA, i = idx % N_LAMBDA;
  phi[idx] = rexp<Real>(-lambda2_of<Real>(i) * T_ratio[v]);   // line_shape_function, los_tracker.hpp:142-147
}

constexpr int LANES_PER_RAY = 4;
constexpr int RAYS_PER_WARP = 32 / LANES_PER_RAY;
constexpr int LAMBDA_PER_LANE = N_LAMBDA / LANES_PER_RAY;   // 5

// MODE 0: rows of the influence matrix (voxel-origin rays); MODE 1: sun-ward rays -> S0, tau
template <class Real, int MODE>
__global__ void __launch_bounds__(256)
march_kernel(GridView<Real> g, EmissionView<Real> em, int v_begin, long long n_rays_total,
             ListView<Real> lists, const int *__restrict__ shadow, double *__restrict__ K,
             double *__restrict__ S0, double *__restrict__ tau_sp_out, double *__restrict__ tau_abs_out,
             int *work_counter, unsigned long long *step_counter) {
  const int lane = threadIdx.x & 31;
  const int sub = lane & (LANES_PER_RAY - 1);
  const int grp = lane / LANES_PER_RAY;
  const int cap = lists.cap;
  const long long n_tasks = (n_rays_total + RAYS_PER_WARP - 1) / RAYS_PER_WARP;

  Real w[LAMBDA_PER_LANE], l2[LAMBDA_PER_LANE];
#pragma unroll
  for (int m = 0; m < LAMBDA_PER_LANE; m++) {
    w[m] = weight_of<Real>(sub + LANES_PER_RAY * m);
    l2[m] = lambda2_of<Real>(sub + LANES_PER_RAY * m);
  }

  unsigned long long my_steps = 0;
  while (true) {
    long long task = 0;
    if (lane == 0) task = atomicAdd(work_counter, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;

    const long long ray = task * RAYS_PER_WARP + grp;
    const bool valid = ray < n_rays_total;
    int len = 0, v0 = 0, ir = 0;
    if (valid) {
      len = lists.len[ray];
      if (MODE == 0) { v0 = v_begin + (int) (ray / g.n_rays); ir = (int) (ray % g.n_rays); }
      else { v0 = (int) ray; if (shadow[v0]) len = -1; }
    }
    const Real Tr0 = valid ? em.T_ratio[v0] : Real(1);
    const Real renorm = one_over_sqrt_pi<Real>() * rsqrt_<Real>(Tr0);   // line_shape_normalization
    Real ls0[LAMBDA_PER_LANE], P[LAMBDA_PER_LANE];
#pragma unroll
    for (int m = 0; m < LAMBDA_PER_LANE; m++) {
      ls0[m] = rexp<Real>(-l2[m] * Tr0);
      P[m] = Real(1);
    }
    const Real domega = (MODE == 0 && valid) ? g.ray_domega[ir] : Real(1);
    Real tau_sp = 0, tau_abs = 0;

    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));

    const Real *dl = lists.dist + (size_t) (valid ? ray : 0) * cap;
    const int *el = lists.ent + (size_t) (valid ? ray : 0) * cap;
    Real dprev = (len > 1) ? dl[0] : Real(0);
    int vox = (len > 1) ? el[0] : 0;

    for (int k = 1; k < maxlen; k++) {
      const bool active = k < len;
      Real G = 0;
      int vox_next = 0;
      if (active) {
#ifdef B200RT_CHECKS
        if (vox < 0 || vox >= g.n_vox || k >= cap || v0 < 0 || v0 >= g.n_vox)
          printf("march<%d>: bad index ray=%lld k=%d len=%d vox=%d v0=%d cap=%d\n", MODE, ray, k, len, vox, v0, cap);
#endif
        const Real dk = dl[k];
        vox_next = el[k];
        const Real s = dk - dprev;                                   // boundaries.hpp:366,376
        dprev = dk;
        const Real dts = em.dtau_species[vox];
        const Real dta = em.dtau_absorber[vox];
        if (MODE == 1) { tau_sp += dts * s; tau_abs += dta * s; }
        const Real *ph = em.phi + (size_t) vox * N_LAMBDA + sub;
#pragma unroll
        for (int m = 0; m < LAMBDA_PER_LANE; m++) {
          const Real lineshape = ph[LANES_PER_RAY * m];
          const Real tau = (dta + dts * lineshape) * s;
          const Real tp = rexp<Real>(-tau);
          const Real Pf = P[m] * tp;
          if (MODE == 0) {
            Real c = ((double) tau < 1e-3) ? (Real(1.0) - Real(0.5) * tau) : (Real(1.0) - tp) / tau;
            c *= (w[m] * lineshape * P[m] * dts * s);
            G += c * renorm * ls0[m];
          }
          P[m] = Pf;
        }
      }
      if (MODE == 0) {
        G += __shfl_xor_sync(0xffffffffu, G, 1);
        G += __shfl_xor_sync(0xffffffffu, G, 2);
        if (active && sub == 0) atomicAdd(&K[(size_t) v0 * g.n_vox + vox], (double) (domega * G));
      }
      vox = vox_next;
    }
    if (MODE == 0) {
      if (sub == 0 && len > 1) my_steps += (unsigned long long) (len - 1);
    } else {
      // holstein_T_final = sum_i w_i * norm(T0) * phi0_i * P_i   (singlet_CFR.hpp:183-186)
      Real T = 0;
#pragma unroll
      for (int m = 0; m < LAMBDA_PER_LANE; m++) T += w[m] * renorm * ls0[m] * P[m];
      T += __shfl_xor_sync(0xffffffffu, T, 1);
      T += __shfl_xor_sync(0xffffffffu, T, 2);
      if (valid && sub == 0) {
        if (len == -1) { S0[v0] = 0.0; tau_sp_out[v0] = -1.0; tau_abs_out[v0] = -1.0; }   // behind the limb
        else if (len <= 1) { S0[v0] = 1.0; tau_sp_out[v0] = 0.0; tau_abs_out[v0] = 0.0; } // tracker reset values
        else { S0[v0] = (double) T; tau_sp_out[v0] = (double) tau_sp; tau_abs_out[v0] = (double) tau_abs; }
      }
    }
  }
  if (MODE == 0 && step_counter) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_steps += __shfl_xor_sync(0xffffffffu, my_steps, o);
    if (lane == 0 && my_steps) atomicAdd(step_counter, my_steps);
  }
}

} // namespace

template <class Real>
cudaError_t launch_phi_table(const Real *T_ratio, int n_vox, Real *phi, cudaStream_t s) {
  const int n = n_vox * N_LAMBDA;
  phi_table_kernel<Real><<<(n + 255) / 256, 256, 0, s>>>(T_ratio, n_vox, phi);
  return cudaGetLastError();
}

template <class Real>
cudaError_t launch_influence(const GridView<Real> &g, const EmissionView<Real> &em, int v_begin, int v_end,
                             ListView<Real> lists, double *K, int *work_counter,
                             unsigned long long *step_counter, cudaStream_t s) {
  const long long n = (long long) (v_end - v_begin) * g.n_rays;
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  const int threads = 256;
  const long long tasks = (n + RAYS_PER_WARP - 1) / RAYS_PER_WARP;
  long long blocks = (tasks + threads / 32 - 1) / (threads / 32);
  const long long persistent = (long long) NUM_SMS * 4;
  if (blocks > persistent) blocks = persistent;
  march_kernel<Real, 0><<<(unsigned) blocks, threads, 0, s>>>(g, em, v_begin, n, lists, nullptr, K, nullptr, nullptr,
                                                              nullptr, work_counter, step_counter);
  return cudaGetLastError();
}

template <class Real>
cudaError_t launch_single_scattering(const GridView<Real> &g, const EmissionView<Real> &em, ListView<Real> lists,
                                     const int *shadow, double *S0, double *tau_sp, double *tau_abs,
                                     int *work_counter, cudaStream_t s) {
  const long long n = g.n_vox;
  cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  const int threads = 256;
  const long long tasks = (n + RAYS_PER_WARP - 1) / RAYS_PER_WARP;
  long long blocks = (tasks + threads / 32 - 1) / (threads / 32);
  if (blocks > NUM_SMS * 4) blocks = NUM_SMS * 4;
  march_kernel<Real, 1><<<(unsigned) blocks, threads, 0, s>>>(g, em, 0, n, lists, shadow, nullptr, S0, tau_sp, tau_abs,
                                                              work_counter, nullptr);
  return cudaGetLastError();
}

template cudaError_t launch_phi_table<double>(const double *, int, double *, cudaStream_t);
template cudaError_t launch_phi_table<float>(const float *, int, float *, cudaStream_t);
template cudaError_t launch_influence<double>(const GridView<double> &, const EmissionView<double> &, int, int,
                                              ListView<double>, double *, int *, unsigned long long *, cudaStream_t);
template cudaError_t launch_influence<float>(const GridView<float> &, const EmissionView<float> &, int, int,
                                             ListView<float>, double *, int *, unsigned long long *, cudaStream_t);
template cudaError_t launch_single_scattering<double>(const GridView<double> &, const EmissionView<double> &,
                                                      ListView<double>, const int *, double *, double *, double *,
                                                      int *, cudaStream_t);
template cudaError_t launch_single_scattering<float>(const GridView<float> &, const EmissionView<float> &,
                                                     ListView<float>, const int *, double *, double *, double *,
                                                     int *, cudaStream_t);

} // namespace b200rt
