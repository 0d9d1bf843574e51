// solve.cu -- dense source-function solve (I - w K) S = S0 on the device (sm_100a).
//
// Replaces emission_voxels::solve (Eigen PartialPivLU, reference
// emission/emission_voxels.hpp:170-176) + singlet_CFR::pre_solve (:402-404) and the
// reference GPU path's cuSOLVER Sgetrf/Sgetrs + 2x cublasSgeam + prepare kernel
// (emission_voxels.hpp:307-478, singlet_CFR.hpp:641-668).  Always FP64.
//
// Algorithm: right-looking blocked LU, block size 64, of the row-major matrix
// A = I - w K.  Every row of K is a set of scattering probabilities (entries >= 0,
// row sum < 1), so A is strictly diagonally dominant by rows; for such matrices
// Gaussian elimination needs no row exchanges (growth factor <= 2), and partial
// pivoting applied to A^T would provably pick the diagonal every time.  The
// dominance margin is checked on the device while A is formed and the call fails
// with B200RT_ERR_NOT_DOMINANT if it does not hold; the residual of the returned
// solution is computed in FP64 and reported.
//
// Per block step k:  (1) diag_kernel   : LU of the 64x64 diagonal block in shared memory and
//                                         explicit inverses of its two triangular factors;
//                    (2) panel_kernel  : L21 = A21 * U11^-1,  U12 = L11^-1 * A12,  y_k = L11^-1 b_k
//                                         as 64x64x64 products on the FP64 tensor pipe;
//                    (3) update_kernel : A22 -= L21 * U12 (128x128 tiles, k = 64), b2 -= L21 y_k.
// All three use mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4: tcgen05 has no FP64 kind, so
// this is the FP64 tensor-core path on sm_100a).  Back substitution walks the block
// columns from the right with the stored U_kk^-1.
#include <algorithm>
#include <cmath>
#include <vector>
#include "common.hpp"

namespace b200rt {

namespace {

constexpr int NB = 64;

__device__ __forceinline__ void dmma8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- A = I - w K (padded to np, identity in the padding), b = S0, row margins
__global__ void prepare_kernel(const double *__restrict__ K, int n, int np, double w, const double *__restrict__ S0,
                               double *__restrict__ A, double *__restrict__ b, double *__restrict__ margin) {
  const int i = blockIdx.x;
  __shared__ double red[256];
  double off = 0, diag = 0;
  double *Ai = A + (size_t) i * np;
  if (i < n) {
    const double *Ki = K + (size_t) i * n;
    for (int j = threadIdx.x; j < np; j += blockDim.x) {
      double v = 0;
      if (j < n) {
        v = -w * Ki[j];
        if (j == i) { v += 1.0; diag = v; }
        else off += fabs(v);
      }
      Ai[j] = v;
    }
  } else {
    for (int j = threadIdx.x; j < np; j += blockDim.x) Ai[j] = (j == i) ? 1.0 : 0.0;
    if (threadIdx.x == 0) diag = 1.0;
  }
  red[threadIdx.x] = fabs(diag) - off;   // partial: sum over threads gives |a_ii| - sum|a_ij|
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    margin[i] = red[0];
    b[i] = (i < n) ? S0[i] : 0.0;
  }
}

// ---- (1) diagonal block: LU without exchanges + inverses of L (unit lower) and U
__global__ void __launch_bounds__(256)
diag_kernel(double *__restrict__ A, int np, int k, double *__restrict__ dinv) {
  extern __shared__ __align__(16) double dsm[];
  double (*a)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm);
  double (*li)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + NB * (NB + 1));
  double (*ui)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(dsm + 2 * NB * (NB + 1));
  double *Akk = A + ((size_t) k * NB) * np + (size_t) k * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    a[r][c] = Akk[(size_t) r * np + c];
  }
  __syncthreads();
  for (int j = 0; j < NB - 1; j++) {
    const double inv = 1.0 / a[j][j];
    __syncthreads();
    if (threadIdx.x > j && threadIdx.x < NB) a[threadIdx.x][j] *= inv;
    __syncthreads();
    const int m = NB - 1 - j;
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int r = j + 1 + e / m, c = j + 1 + e % m;
      a[r][c] -= a[r][j] * a[j][c];
    }
    __syncthreads();
  }
  // inverses, one column per thread: threads 0..63 -> L^-1, threads 64..127 -> U^-1
  if (threadIdx.x < NB) {
    const int c = threadIdx.x;
    for (int i = 0; i < NB; i++) {
      double s = (i == c) ? 1.0 : 0.0;
      if (i > c) { for (int j = c; j < i; j++) s -= a[i][j] * li[j][c]; }
      li[i][c] = (i < c) ? 0.0 : s;
    }
  } else if (threadIdx.x < 2 * NB) {
    const int c = threadIdx.x - NB;
    for (int i = NB - 1; i >= 0; i--) {
      double s = (i == c) ? 1.0 : 0.0;
      if (i < c) { for (int j = i + 1; j <= c; j++) s -= a[i][j] * ui[j][c]; }
      ui[i][c] = (i > c) ? 0.0 : s / a[i][i];
    }
  }
  __syncthreads();
  double *Li = dinv + (size_t) k * 2 * NB * NB, *Ui = Li + NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    Akk[(size_t) r * np + c] = a[r][c];
    Li[e] = li[r][c];
    Ui[e] = ui[r][c];
  }
}

// 64x64x64 product on the tensor pipe: out = P * Q, both staged in shared memory
// (P row stride SP, Q row stride SQ chosen so the fragment loads are conflict free).
constexpr int SP = NB + 4;   // 68: (row*68 + col) mod 16 distinct for row,col < 4
constexpr int SQ = NB + 4;
__device__ __forceinline__ void tile_product_64(const double *Ps, const double *Qs, double (&acc)[4][4][2], int warp,
                                                int lane) {
  // 4 warps: warp tile 32x32 = 4 (m) x 4 (n) DMMA tiles
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll 4
  for (int k0 = 0; k0 < NB; k0 += 4) {
    double af[4], bf[4];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) af[mt] = Ps[(wm + 8 * mt + g) * SP + k0 + tq];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) bf[nt] = Qs[(k0 + tq) * SQ + wn + 8 * nt + g];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
}

// ---- (2) panel: blockIdx.x < nrest : L21 tile (rows below) = A * Uinv
//                 else               : U12 tile (cols right)  = Linv * A ; last block: y_k = Linv b_k
__global__ void __launch_bounds__(128)
panel_kernel(double *__restrict__ A, int np, int k, const double *__restrict__ dinv, double *__restrict__ b) {
  extern __shared__ __align__(16) double sm[];
  double *Ps = sm, *Qs = sm + NB * SP;
  const int nrest = np / NB - k - 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double *Li = dinv + (size_t) k * 2 * NB * NB, *Ui = Li + NB * NB;
  const int bid = blockIdx.x;
  if (bid == 2 * nrest) {   // y_k = Linv * b_k
    double *bk = b + (size_t) k * NB;
    __shared__ double bs[NB];
    if (threadIdx.x < NB) bs[threadIdx.x] = bk[threadIdx.x];
    __syncthreads();
    if (threadIdx.x < NB) {
      double s = 0;
      for (int j = 0; j <= (int) threadIdx.x; j++) s += Li[threadIdx.x * NB + j] * bs[j];
      bk[threadIdx.x] = s;
    }
    return;
  }
  const bool lower = bid < nrest;
  const int t = lower ? bid : bid - nrest;
  double *tile = lower ? A + ((size_t) (k + 1 + t) * NB) * np + (size_t) k * NB
                       : A + ((size_t) k * NB) * np + (size_t) (k + 1 + t) * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    const double tv = tile[(size_t) r * np + c];
    if (lower) { Ps[r * SP + c] = tv; Qs[r * SQ + c] = Ui[e]; }
    else { Ps[r * SP + c] = Li[e]; Qs[r * SQ + c] = tv; }
  }
  __syncthreads();
  double acc[4][4][2];
  tile_product_64(Ps, Qs, acc, warp, lane);
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32, g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      double *o = tile + (size_t) (wm + 8 * mt + g) * np + wn + 8 * nt + 2 * tq;
      *reinterpret_cast<double2 *>(o) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// ---- (3) trailing update: C[128x128 tile] -= L21[128x64] * U12[64x128];  b2 -= L21 * y_k
constexpr int TM = 128, TN = 128;
constexpr int SA = NB + 4;    // 68
constexpr int SB = TN + 4;    // 132: (k*132 + n) mod 16 distinct for k,n < 4
__global__ void __launch_bounds__(256)
update_kernel(double *__restrict__ A, int np, int k, double *__restrict__ b) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm;               // [TM][SA]
  double *Bs = sm + TM * SA;     // [NB][SB]
  const int row0 = (k + 1) * NB + blockIdx.y * TM;
  const int col0 = (k + 1) * NB + blockIdx.x * TN;
  const int kc = k * NB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // stage L21 rows [row0, row0+128) x cols [kc, kc+64)
  for (int e = threadIdx.x; e < TM * (NB / 2); e += blockDim.x) {
    const int r = e / (NB / 2), c2 = e % (NB / 2);
    double2 v = make_double2(0.0, 0.0);
    if (row0 + r < np) v = *reinterpret_cast<const double2 *>(A + (size_t) (row0 + r) * np + kc + 2 * c2);
    *reinterpret_cast<double2 *>(As + r * SA + 2 * c2) = v;
  }
  // stage U12 rows [kc, kc+64) x cols [col0, col0+128)
  for (int e = threadIdx.x; e < NB * (TN / 2); e += blockDim.x) {
    const int r = e / (TN / 2), c2 = e % (TN / 2);
    double2 v = make_double2(0.0, 0.0);
    if (col0 + 2 * c2 < np) v = *reinterpret_cast<const double2 *>(A + (size_t) (kc + r) * np + col0 + 2 * c2);
    *reinterpret_cast<double2 *>(Bs + r * SB + 2 * c2) = v;
  }
  __syncthreads();

  // 8 warps: 4 (m) x 2 (n); warp tile 32 x 64 = 4 x 8 DMMA tiles
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
  const int g = lane >> 2, tq = lane & 3;
  double acc[4][8][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#pragma unroll 2
  for (int k0 = 0; k0 < NB; k0 += 4) {
    double af[4], bf[8];
#pragma unroll
    for (int mt = 0; mt < 4; mt++) af[mt] = As[(wm + 8 * mt + g) * SA + k0 + tq];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) bf[nt] = Bs[(k0 + tq) * SB + wn + 8 * nt + g];
#pragma unroll
    for (int mt = 0; mt < 4; mt++)
#pragma unroll
      for (int nt = 0; nt < 8; nt++) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++) {
    const int r = row0 + wm + 8 * mt + g;
    if (r < np) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        const int c = col0 + wn + 8 * nt + 2 * tq;
        if (c < np) {
          double2 *p = reinterpret_cast<double2 *>(A + (size_t) r * np + c);
          double2 v = *p;
          v.x -= acc[mt][nt][0];
          v.y -= acc[mt][nt][1];
          *p = v;
        }
      }
    }
  }
  // right-hand side: b[row] -= L21[row][:] . y_k   (first tile column only)
  if (blockIdx.x == 0 && threadIdx.x < TM) {
    const int r = row0 + threadIdx.x;
    if (r < np) {
      const double *yk = b + kc;
      double s = 0;
      for (int j = 0; j < NB; j++) s += As[threadIdx.x * SA + j] * yk[j];
      b[r] -= s;
    }
  }
}

// ---- back substitution, block column k: x_k = Uinv_kk y_k ; y_i -= U_ik x_k for i < k
__global__ void __launch_bounds__(64)
backsolve_kernel(const double *__restrict__ A, int np, int k, const double *__restrict__ dinv, double *__restrict__ b,
                 double *__restrict__ x) {
  __shared__ double xk[NB], yk[NB];
  const double *Ui = dinv + (size_t) k * 2 * NB * NB + NB * NB;
  const int tid = threadIdx.x;
  yk[tid] = b[(size_t) k * NB + tid];
  __syncthreads();
  double s = 0;
  for (int j = tid; j < NB; j++) s += Ui[tid * NB + j] * yk[j];
  xk[tid] = s;
  __syncthreads();
  const int i = blockIdx.x;   // block row i < k updates, block k writes the solution
  if (i == k) { x[(size_t) k * NB + tid] = xk[tid]; return; }
  const double *Uik = A + ((size_t) i * NB + tid) * np + (size_t) k * NB;
  double u = 0;
  for (int j = 0; j < NB; j++) u += Uik[j] * xk[j];
  b[(size_t) i * NB + tid] -= u;
}

// ---- residual of the ORIGINAL system: |S0 - (S - w K S)|, one warp per row
__global__ void residual_kernel(const double *__restrict__ K, int n, double w, const double *__restrict__ S0,
                                const double *__restrict__ S, double *__restrict__ rabs) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double *Kr = K + (size_t) row * n;
  double s = 0;
  for (int j = lane; j < n; j += 32) s += Kr[j] * S[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) rabs[row] = fabs(S0[row] - (S[row] - w * s));
}

template <class Real>
__global__ void convert_kernel(const double *__restrict__ src, Real *__restrict__ dst, long long n) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (Real) src[i];
}

} // namespace

template <class Real>
cudaError_t launch_convert(const double *src, Real *dst, long long n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  convert_kernel<Real><<<(unsigned) ((n + 255) / 256), 256, 0, s>>>(src, dst, n);
  return cudaGetLastError();
}
template cudaError_t launch_convert<double>(const double *, double *, long long, cudaStream_t);
template cudaError_t launch_convert<float>(const double *, float *, long long, cudaStream_t);

int solve_dense(b200rt_ctx *c, int n, const double *K, double branching, const double *S0, double *S,
                SolveResult *res) {
  const int np = ((n + 127) / 128) * 128;
  const int nblk = np / NB;
  cudaStream_t st = c->stream;
  // workspace: A[np*np] | b[np] | x[np] | margin[np] | rabs[np]
  B200RT_CUDA(c, c->lu.ensure(((size_t) np * np + 4 * (size_t) np) * sizeof(double)));
  B200RT_CUDA(c, c->lu_dinv.ensure((size_t) nblk * 2 * NB * NB * sizeof(double)));
  double *A = c->lu.as<double>();
  double *b = A + (size_t) np * np, *x = b + np, *margin = x + np, *rabs = margin + np;
  double *dinv = c->lu_dinv.as<double>();
  int launches = 0;

  prepare_kernel<<<np, 256, 0, st>>>(K, n, np, branching, S0, A, b, margin);
  launches++;
  B200RT_CUDA(c, cudaGetLastError());
  std::vector<double> hm(np);
  B200RT_CUDA(c, cudaMemcpyAsync(hm.data(), margin, np * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaStreamSynchronize(st));
  double min_margin = 1e300;
  for (int i = 0; i < n; i++) min_margin = std::min(min_margin, hm[i]);
  if (res) res->min_margin = min_margin;
  if (!(min_margin > 0.0))
    return fail(c, B200RT_ERR_NOT_DOMINANT,
                "I - w*K is not strictly row diagonally dominant (min margin " + std::to_string(min_margin) +
                    "): the influence matrix rows are not scattering probabilities");

  const size_t panel_smem = (size_t) (NB * SP + NB * SQ) * sizeof(double);
  const size_t update_smem = (size_t) (TM * SA + NB * SB) * sizeof(double);
  const size_t diag_smem = (size_t) 3 * NB * (NB + 1) * sizeof(double);
  B200RT_CUDA(c, cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) diag_smem));
  B200RT_CUDA(c, cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) panel_smem));
  B200RT_CUDA(c, cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) update_smem));

  for (int k = 0; k < nblk; k++) {
    diag_kernel<<<1, 256, diag_smem, st>>>(A, np, k, dinv);
    const int nrest = nblk - k - 1;
    panel_kernel<<<2 * nrest + 1, 128, panel_smem, st>>>(A, np, k, dinv, b);
    launches += 2;
    if (nrest > 0) {
      const int M = nrest * NB;
      dim3 grid((M + TN - 1) / TN, (M + TM - 1) / TM);
      update_kernel<<<grid, 256, update_smem, st>>>(A, np, k, b);
      launches++;
    }
  }
  for (int k = nblk - 1; k >= 0; k--) {
    backsolve_kernel<<<k + 1, NB, 0, st>>>(A, np, k, dinv, b, x);
    launches++;
  }
  B200RT_CUDA(c, cudaGetLastError());
  B200RT_CUDA(c, cudaMemcpyAsync(S, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  residual_kernel<<<(n + 7) / 8, 256, 0, st>>>(K, n, branching, S0, S, rabs);
  launches++;
  B200RT_CUDA(c, cudaGetLastError());
  std::vector<double> hr(n), hs0(n);
  B200RT_CUDA(c, cudaMemcpyAsync(hr.data(), rabs, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaMemcpyAsync(hs0.data(), S0, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaStreamSynchronize(st));
  double rmax = 0, smax = 0;
  for (int i = 0; i < n; i++) { rmax = std::max(rmax, hr[i]); smax = std::max(smax, std::fabs(hs0[i])); }
  if (res) { res->residual = rmax / (smax > 0 ? smax : 1.0); res->launches = launches; }
  return B200RT_OK;
}

} // namespace b200rt
