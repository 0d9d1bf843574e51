// solve.cu -- dense source-function solve (I - w K) S = S0 on the device (sm_100a).
//
// Replaces emission_voxels::solve (Eigen PartialPivLU, reference
// emission/emission_voxels.hpp:170-176) + singlet_CFR::pre_solve (:402-404) and the
// reference GPU path's cuSOLVER Sgetrf/Sgetrs + 2x cublasSgeam + prepare kernel
// (emission_voxels.hpp:307-478, singlet_CFR.hpp:641-668).  Always FP64.
//
// Algorithm: right-looking BLOCK LU (block size 128) of the row-major matrix A = I - w K with
// explicitly inverted diagonal blocks:
//     A = [ I    0 ] [ A11  A12 ]        L21 = A21 * A11^-1,   S = A22 - L21 * A12,
//         [ L21  I ] [ 0    S   ]
// Every row of K is a set of scattering probabilities (entries >= 0, row sum < 1), so A is strictly
// diagonally dominant by rows, and so is every Schur complement S: no row exchanges are needed
// (partial pivoting would pick the diagonal every time) and the diagonal blocks are safely
// invertible by Gauss-Jordan without pivoting.  The dominance margin is checked on the device
// while A is formed and the call fails with B200RT_ERR_NOT_DOMINANT if it does not hold; the
// residual of the returned solution is computed in FP64 and reported.
//
// Per block column KB, the "chain" (critical path, on a high-priority stream):
//   gj128_kernel   : in-register Gauss-Jordan inverse of the 128x128 diagonal block (one CTA, each
//                    thread owns an 8x8 tile, pivot row / column broadcast through shared memory);
//   gemm128<PANEL> : L21 = A21 * A11^-1 (128x128 tiles, DMMA);
//   rhs128_kernel  : b2 -= L21 b1;
// then gemm128<UPDATE>: A22 -= L21 * A12 -- 128x128 tiles, k = 128, 4-stage cp.async pipeline,
//                    accumulators initialised from the C tile (one read and one write of A22 per
//                    block column).
// Look-ahead: the update of block column KB is split into the L-shaped part next to the diagonal
// (what chain KB+1 needs) and the rest; chain KB+1 runs concurrently with the rest.
// All products use mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4: tcgen05 has no FP64 kind, so
// this is the FP64 tensor-core path on sm_100a).  Back substitution walks the block columns from
// the right:  x_K = A_KK^-1 y_K ;  y_I -= A_IK x_K  (I < K).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "common.hpp"

namespace b200rt {

namespace {


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

// ---- (1) diagonal block: in-place Gauss-Jordan inverse, no pivoting.
// 256 threads = 16 x 16, thread (ty, tx) owns rows ty*8.., columns tx*8.. in registers.  The inner
// 8 steps are unrolled so that every register index is static.
__device__ __forceinline__ double rcp_fast(double b) {   // 1/b for normal b, <= 1 ulp
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  return fma(r, e, r);
}
constexpr int TB = 128;   // block size of the factorisation
__global__ void __launch_bounds__(256)
gj128_kernel(const double *__restrict__ A, int np, int KB, double *__restrict__ dinv) {
  __shared__ __align__(16) double prow[2][TB], pcol[2][TB];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const double *Akk = A + ((size_t) KB * TB) * np + (size_t) KB * TB;
  double a[8][8];
#pragma unroll
  for (int rr = 0; rr < 8; rr++)
#pragma unroll
    for (int c2 = 0; c2 < 4; c2++) {
      const double2 v = *reinterpret_cast<const double2 *>(Akk + (size_t) (ty * 8 + rr) * np + tx * 8 + 2 * c2);
      a[rr][2 * c2] = v.x;
      a[rr][2 * c2 + 1] = v.y;
    }
  for (int jb = 0; jb < 16; jb++) {
#pragma unroll
    for (int jj = 0; jj < 8; jj++) {
      const int j = jb * 8 + jj, buf = jj & 1;
      if (ty == jb) {
#pragma unroll
        for (int cc = 0; cc < 8; cc++) prow[buf][tx * 8 + cc] = a[jj][cc];
      }
      if (tx == jb) {
#pragma unroll
        for (int rr = 0; rr < 8; rr++) pcol[buf][ty * 8 + rr] = a[rr][jj];
      }
      __syncthreads();
      const double p = rcp_fast(prow[buf][j]);
      double sp[8], f[8];
#pragma unroll
      for (int cc = 0; cc < 8; cc++) sp[cc] = prow[buf][tx * 8 + cc] * p;
#pragma unroll
      for (int rr = 0; rr < 8; rr++) f[rr] = pcol[buf][ty * 8 + rr];
#pragma unroll
      for (int rr = 0; rr < 8; rr++)
#pragma unroll
        for (int cc = 0; cc < 8; cc++) a[rr][cc] = fma(-f[rr], sp[cc], a[rr][cc]);
      if (ty == jb) {
#pragma unroll
        for (int cc = 0; cc < 8; cc++) a[jj][cc] = sp[cc];
      }
      if (tx == jb) {
#pragma unroll
        for (int rr = 0; rr < 8; rr++) a[rr][jj] = -f[rr] * p;
        if (ty == jb) a[jj][jj] = p;
      }
    }
  }
  double *out = dinv + (size_t) KB * TB * TB;
#pragma unroll
  for (int rr = 0; rr < 8; rr++)
#pragma unroll
    for (int c2 = 0; c2 < 4; c2++)
      *reinterpret_cast<double2 *>(out + (size_t) (ty * 8 + rr) * TB + tx * 8 + 2 * c2) =
          make_double2(a[rr][2 * c2], a[rr][2 * c2 + 1]);
}

// ---- rhs: b[r] -= L21[r][block column KB] . b_KB for every row below (one warp per row)
__global__ void __launch_bounds__(256)
rhs128_kernel(const double *__restrict__ Lbuf, int np, int KB, double *__restrict__ b) {
  const int row = (KB + 1) * TB + blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= np) return;
  const double *Ar = Lbuf + (size_t) row * TB;
  const double *yk = b + (size_t) KB * TB;
  double s = 0;
#pragma unroll
  for (int q = 0; q < TB / 32; q++) s += Ar[lane + 32 * q] * yk[lane + 32 * q];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) b[row] -= s;
}

// ---- 128x128x128 tile products on the FP64 tensor pipe
constexpr int TM = 128, TN = 128, TK = 128;
constexpr int KC = 16;                 // k-chunk per pipeline stage
constexpr int STAGES = 4;
constexpr int SA = KC + 4;             // 20: (row*20 + col) mod 16 distinct for row, col < 4  (conflict-free LDS.64)

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// MODE 0: update, the L-shaped set of tiles next to the diagonal (block column KB+1 from the diagonal
//         down, block row KB+1 right of the diagonal) -- what the next chain needs:  C -= L21 * A12
// MODE 1: update, the square (i, j >= KB+2)
// MODE 2: panel,  tile (KB+1+blockIdx.x, KB):  L21 = A21 * A_KK^-1 -> Lbuf[row][0..127] (the L panels live in
//         a double-buffered side array: the updates of block column KB read one half while chain KB+1 fills the other)
// NT = DMMA n-tiles per warp: 8 -> 128x128 output tile per CTA (bulk update), 4 -> 128x64 (the two
// kernels on the critical path run as twice as many CTAs with half the latency; blockIdx.y picks the half).
template <int MODE, int NT>
__global__ void __launch_bounds__(256)
gemm128_kernel(double *__restrict__ A, int np, int KB, const double *__restrict__ dinv, double *__restrict__ Lbuf) {
  constexpr int TNc = 16 * NT;           // output tile width of this CTA
  constexpr int SBc = TNc + 4;           // 132 / 68: (k*SB + n) mod 16 distinct for k, n < 4
  constexpr int STAGE = TM * SA + KC * SBc;
  extern __shared__ __align__(16) double sm[];
  const int nB = np / TM;
  int ti, tj;
  if (MODE == 0) {
    const int ncol = nB - KB - 1;
    if ((int) blockIdx.x < ncol) { ti = KB + 1 + blockIdx.x; tj = KB + 1; }
    else { ti = KB + 1; tj = KB + 2 + (blockIdx.x - ncol); }
  } else if (MODE == 1) {
    const int m = nB - KB - 2;
    ti = KB + 2 + blockIdx.x / m;
    tj = KB + 2 + blockIdx.x % m;
  } else {
    ti = KB + 1 + blockIdx.x;
    tj = KB;
  }
  const int row0 = ti * TM, col0 = tj * TN + blockIdx.y * TNc, kc = KB * TK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // L operand [128 rows][128 k]: the panel reads A21 in place, the updates read the panel's output
  const double *Lg = (MODE == 2) ? A + (size_t) row0 * np + kc : Lbuf + (size_t) row0 * TB;
  const int ldl = (MODE == 2) ? np : TB;
  const double *Ug = (MODE == 2) ? dinv + (size_t) KB * TB * TB + blockIdx.y * TNc : A + (size_t) kc * np + col0;
  const int ldu = (MODE == 2) ? TB : np;
  double *Cg = A + (size_t) row0 * np + col0;

  auto issue = [&](int chunk) {
    double *As = sm + (chunk % STAGES) * STAGE;
    double *Bs = As + TM * SA;
    const int k0 = chunk * KC;
#pragma unroll
    for (int i = 0; i < 4; i++) {       // A chunk: 128 rows x 16 doubles = 1024 x 16 B
      const int r = (tid >> 3) + 32 * i, p = tid & 7;
      cp_async16(As + r * SA + 2 * p, Lg + (size_t) r * ldl + k0 + 2 * p);
    }
    constexpr int PPR = TNc / 2;        // 16-byte pieces per B row
#pragma unroll
    for (int i = 0; i < KC * PPR / 256; i++) {
      const int e = tid + 256 * i, kk = e / PPR, p = e % PPR;
      cp_async16(Bs + kk * SBc + 2 * p, Ug + (size_t) (k0 + kk) * ldu + 2 * p);
    }
  };
  constexpr int NCHUNK = TK / KC;
#pragma unroll
  for (int c = 0; c < STAGES - 1; c++) { issue(c); cp_async_commit(); }

  // 8 warps: 4 (m) x 2 (n); warp tile 32 x (8 NT) = 4 x NT DMMA tiles
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * (8 * NT);
  const int g = lane >> 2, tq = lane & 3;
  double acc[4][NT][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      if (MODE == 2) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
      else {   // the accumulators start from C; the A fragments are negated below
        const double2 v = *reinterpret_cast<const double2 *>(Cg + (size_t) (wm + 8 * mt + g) * np + wn + 8 * nt + 2 * tq);
        acc[mt][nt][0] = v.x;
        acc[mt][nt][1] = v.y;
      }
    }

  for (int c = 0; c < NCHUNK; c++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (c + STAGES - 1 < NCHUNK) issue(c + STAGES - 1);
    cp_async_commit();
    const double *As = sm + (c % STAGES) * STAGE;
    const double *Bs = As + TM * SA;
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += 4) {
      double af[4], bf[NT];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) {
        const double v = As[(wm + 8 * mt + g) * SA + k0 + tq];
        af[mt] = (MODE == 2) ? v : -v;
      }
#pragma unroll
      for (int nt = 0; nt < NT; nt++) bf[nt] = Bs[(k0 + tq) * SBc + wn + 8 * nt + g];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  // the panel does not write in place: the CTAs of the two column halves both read the whole A21 tile
  double *Cout = (MODE == 2) ? Lbuf + (size_t) row0 * TB + blockIdx.y * TNc : Cg;
  const int ldc = (MODE == 2) ? TB : np;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
      *reinterpret_cast<double2 *>(Cout + (size_t) (wm + 8 * mt + g) * ldc + wn + 8 * nt + 2 * tq) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
}

// ---- back substitution, block column KB:  x_K = A_KK^-1 y_K ;  y_I -= A_IK x_K for I < K.
// grid = KB+1 blocks of 512 threads (16 warps x 8 rows, coalesced row reads, all loads of a warp in
// flight together); every block forms x_K itself, block KB stores it, block I < KB updates its rows of y.
__global__ void __launch_bounds__(512)
backsolve128_kernel(const double *__restrict__ A, int np, int KB, const double *__restrict__ dinv,
                    double *__restrict__ b, double *__restrict__ x) {
  __shared__ double xk[TB], yk[TB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double *Ai = dinv + (size_t) KB * TB * TB;
  if (tid < TB) yk[tid] = b[(size_t) KB * TB + tid];
  __syncthreads();
  double v[TB / 32], m[8][TB / 32];
#pragma unroll
  for (int q = 0; q < TB / 32; q++) v[q] = yk[lane + 32 * q];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int q = 0; q < TB / 32; q++) m[i][q] = Ai[(size_t) (warp * 8 + i) * TB + lane + 32 * q];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double s = 0;
#pragma unroll
    for (int q = 0; q < TB / 32; q++) s += m[i][q] * v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) xk[warp * 8 + i] = s;
  }
  __syncthreads();
  const int I = blockIdx.x;
  if (I == KB) { if (tid < TB) x[(size_t) KB * TB + tid] = xk[tid]; return; }
#pragma unroll
  for (int q = 0; q < TB / 32; q++) v[q] = xk[lane + 32 * q];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int q = 0; q < TB / 32; q++) m[i][q] = A[((size_t) I * TB + warp * 8 + i) * np + (size_t) KB * TB + lane + 32 * q];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    double s = 0;
#pragma unroll
    for (int q = 0; q < TB / 32; q++) s += m[i][q] * v[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) b[(size_t) I * TB + warp * 8 + i] -= s;
  }
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

// ---- y_out = w K y_in (one warp per row) and the smallest entry of the row: the M-matrix certificate below
__global__ void kmatvec_kernel(const double *__restrict__ K, int n, double w, const double *__restrict__ y_in,
                               double *__restrict__ y_out, double *__restrict__ row_min) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double *Kr = K + (size_t) row * n;
  double s = 0, mn = 0;
  for (int j = lane; j < n; j += 32) {
    const double k = Kr[j];
    s += k * y_in[j];
    mn = fmin(mn, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { y_out[row] = w * s; if (row_min) row_min[row] = mn; }
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
  const int np = ((n + TB - 1) / TB) * TB;
  const int nK = np / TB;
  cudaStream_t st = c->stream;
  // workspace: A[np*np] | b[np] | x[np] | margin[np] | rabs[np]
  B200RT_CUDA(c, c->lu.ensure(((size_t) np * np + 4 * (size_t) np) * sizeof(double)));
  // block inverses [nK][128][128] | L panels [2][np][128]
  B200RT_CUDA(c, c->lu_dinv.ensure(((size_t) nK * TB * TB + 2 * (size_t) np * TB) * sizeof(double)));
  double *A = c->lu.as<double>();
  double *b = A + (size_t) np * np, *x = b + np, *margin = x + np, *rabs = margin + np;
  double *dinv = c->lu_dinv.as<double>();
  double *Lbuf0 = dinv + (size_t) nK * TB * TB;
  auto Lbuf = [&](int KB) { return Lbuf0 + (size_t) (KB & 1) * np * TB; };
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
  if (!(min_margin > 0.0)) {
    // Not row dominant as it stands (the multiplet emissions: K[(v0,iu),(v,ju)] carries the ORIGIN voxel's lower-state
    // density, multiplet_CFR_emission.hpp:264-271, so its rows are probabilities only after a diagonal rescaling).
    // Certificate instead: K >= 0 and max_i (w K)^m 1 < 1 for some m  =>  rho(w K) < 1  =>  I - w K is a nonsingular
    // M-matrix: every Schur complement is again one, all pivots are positive, and elimination without row exchanges
    // is componentwise backward stable (its error bound is invariant under the diagonal rescaling that makes the
    // matrix row dominant).  x / rabs / margin are free at this point and serve as scratch.
    bool certified = false;
    double kmin = 0, ymax = 1e300;
    std::vector<double> hy(n, 1.0);
    B200RT_CUDA(c, cudaMemcpyAsync(x, hy.data(), n * sizeof(double), cudaMemcpyHostToDevice, st));
    double *ya = x, *yb = rabs;
    const int max_iter = 2048;
    for (int it = 0; it < max_iter && !certified; it++) {
      kmatvec_kernel<<<(n + 7) / 8, 256, 0, st>>>(K, n, branching, ya, yb, it == 0 ? margin : nullptr);
      launches++;
      std::swap(ya, yb);
      if (it == 0) {
        B200RT_CUDA(c, cudaMemcpyAsync(hm.data(), margin, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200RT_CUDA(c, cudaStreamSynchronize(st));
        for (int i = 0; i < n; i++) kmin = std::min(kmin, hm[i]);
        if (kmin < 0.0 || branching < 0.0) break;
      }
      if ((it & 7) == 7 || it == 0) {
        B200RT_CUDA(c, cudaMemcpyAsync(hy.data(), ya, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200RT_CUDA(c, cudaStreamSynchronize(st));
        ymax = 0;
        for (int i = 0; i < n; i++) ymax = std::max(ymax, hy[i]);
        if (ymax < 1.0) certified = true;
        if (!(ymax < 1e200)) break;    // diverging
      }
    }
    B200RT_CUDA(c, cudaGetLastError());
    if (!certified)
      return fail(c, B200RT_ERR_NOT_DOMINANT,
                  "I - w*K is neither strictly row diagonally dominant (min margin " + std::to_string(min_margin) +
                      ") nor certifiably an M-matrix (min K entry " + std::to_string(kmin) + ", max (wK)^m 1 = " +
                      std::to_string(ymax) + "): elimination without row exchanges is not safe");
  }

  const size_t gemm_smem = (size_t) STAGES * (TM * SA + KC * (TN + 4)) * sizeof(double);        // 128x128 tiles
  const size_t gemm_smem_h = (size_t) STAGES * (TM * SA + KC * (TN / 2 + 4)) * sizeof(double);  // 128x64 tiles
  B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_h));
  B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem));
  B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_h));

  // second stream + events for the look-ahead
  if (!c->stream2) {   // the chain is the critical path: give its stream the highest priority
    int lo = 0, hi = 0;
    B200RT_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200RT_CUDA(c, cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
  }
  while ((int) c->lu_events.size() < 2 * nK + 1) {
    cudaEvent_t ev;
    B200RT_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->lu_events.push_back(ev);
  }
  cudaStream_t sB = c->stream2;
  // B200RT_SOLVE_TRACE=1: timestamp every chain / update of the factorisation (development aid)
  static const bool trace = getenv("B200RT_SOLVE_TRACE") != nullptr;
  struct Mark { const char *what; int KB; cudaEvent_t ev; };
  std::vector<Mark> marks;
  auto mark = [&](const char *what, int KB, cudaStream_t s) {
    if (!trace) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, s);
    marks.push_back({what, KB, ev});
  };
  auto chain = [&](int KB, cudaStream_t s) {       // invert the diagonal block, form L21, forward-substitute
    const int m1 = nK - KB - 1;
    gj128_kernel<<<1, 256, 0, s>>>(A, np, KB, dinv);
    launches++;
    if (m1 > 0) {
      gemm128_kernel<2, 4><<<dim3(m1, 2), 256, gemm_smem_h, s>>>(A, np, KB, dinv, Lbuf(KB));
      rhs128_kernel<<<(m1 * TB + 7) / 8, 256, 0, s>>>(Lbuf(KB), np, KB, b);
      launches += 2;
    }
  };
  cudaEvent_t ev_start = c->lu_events[2 * nK];
  B200RT_CUDA(c, cudaEventRecord(ev_start, st));
  B200RT_CUDA(c, cudaStreamWaitEvent(sB, ev_start, 0));
  mark("start", 0, st);
  mark("chain_begin", 0, sB);
  chain(0, sB);
  mark("chain_end", 0, sB);
  B200RT_CUDA(c, cudaEventRecord(c->lu_events[0], sB));            // evP[0]
  for (int KB = 0; KB < nK; KB++) {
    B200RT_CUDA(c, cudaStreamWaitEvent(st, c->lu_events[2 * KB], 0)); // block column KB factored
    const int m1 = nK - KB - 1;                                       // block columns after KB
    if (m1 > 0) {
      mark("ui_begin", KB, st);
      gemm128_kernel<0, 4><<<dim3(2 * m1 - 1, 2), 256, gemm_smem_h, st>>>(A, np, KB, dinv, Lbuf(KB));
      mark("ui_end", KB, st);
      launches++;
      B200RT_CUDA(c, cudaEventRecord(c->lu_events[2 * KB + 1], st));  // evU[KB]
      B200RT_CUDA(c, cudaStreamWaitEvent(sB, c->lu_events[2 * KB + 1], 0));
      mark("chain_begin", KB + 1, sB);
      chain(KB + 1, sB);
      mark("chain_end", KB + 1, sB);
      B200RT_CUDA(c, cudaEventRecord(c->lu_events[2 * KB + 2], sB));  // evP[KB+1]
      const int m2 = m1 - 1;
      if (m2 > 0) {
        gemm128_kernel<1, 8><<<m2 * m2, 256, gemm_smem, st>>>(A, np, KB, dinv, Lbuf(KB));
        mark("uii_end", KB, st);
        launches++;
      }
    }
  }
  mark("factor_end", 0, st);
  for (int KB = nK - 1; KB >= 0; KB--) {
    backsolve128_kernel<<<KB + 1, 512, 0, st>>>(A, np, KB, dinv, b, x);
    launches++;
  }
  mark("backsolve_end", 0, st);
  B200RT_CUDA(c, cudaGetLastError());
  B200RT_CUDA(c, cudaMemcpyAsync(S, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  residual_kernel<<<(n + 7) / 8, 256, 0, st>>>(K, n, branching, S0, S, rabs);
  launches++;
  B200RT_CUDA(c, cudaGetLastError());
  std::vector<double> hr(n), hs0(n);
  B200RT_CUDA(c, cudaMemcpyAsync(hr.data(), rabs, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaMemcpyAsync(hs0.data(), S0, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaStreamSynchronize(st));
  if (trace && !marks.empty()) {
    cudaStreamSynchronize(sB);
    for (auto &m : marks) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[0].ev, m.ev);
      fprintf(stderr, "solve-trace %-12s K=%3d t=%9.3f ms\n", m.what, m.KB, ms);
    }
    for (auto &m : marks) cudaEventDestroy(m.ev);
  }
  double rmax = 0, smax = 0;
  for (int i = 0; i < n; i++) { rmax = std::max(rmax, hr[i]); smax = std::max(smax, std::fabs(hs0[i])); }
  if (res) { res->residual = rmax / (smax > 0 ? smax : 1.0); res->launches = launches; }
  return B200RT_OK;
}

} // namespace b200rt
