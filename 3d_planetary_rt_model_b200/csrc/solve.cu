// solve.cu -- dense source-function solve (I - w K) S = S0 on the device (sm_100a).
//
// Replaces emission_voxels::solve (Eigen PartialPivLU, reference
// emission/emission_voxels.hpp:170-176) + singlet_CFR::pre_solve (:402-404) and the
// reference GPU path's cuSOLVER Sgetrf/Sgetrs + 2x cublasSgeam + prepare kernel
// (emission_voxels.hpp:307-478, singlet_CFR.hpp:641-668).  Always FP64.
//
// Algorithm: right-looking blocked LU of the row-major matrix A = I - w K, two-level blocking
// (64-wide panels inside 128-wide outer blocks).  Every row of K is a set of scattering
// probabilities (entries >= 0, row sum < 1), so A is strictly diagonally dominant by rows; for
// such matrices Gaussian elimination needs no row exchanges (growth factor <= 2), and partial
// pivoting applied to A^T would provably pick the diagonal every time.  The dominance margin is
// checked on the device while A is formed and the call fails with B200RT_ERR_NOT_DOMINANT if it
// does not hold; the residual of the returned solution is computed in FP64 and reported.
//
// Per outer block K (panels k = 2K, 2K+1), the "chain":
//   diag_kernel(k)   : LU of the 64x64 diagonal block in shared memory + explicit inverses of its
//                      two triangular factors;
//   panel_kernel(k)  : L21 = A21 * U11^-1,  U12 = L11^-1 * A12,  y_k = L11^-1 b_k  (64x64x64 DMMA products);
//   rhs_kernel(k)    : b2 -= L21 y_k;
//   strip_kernel(2K) : the 64-wide column / row strips next to panel 2K get A -= L21 U12 so that
//                      panel 2K+1 can be factored;
// then update_kernel : A22 -= [L21(2K) L21(2K+1)] * [U12(2K); U12(2K+1)]  -- 128x128 tiles, k = 128,
//                      4-stage cp.async pipeline, accumulators initialised from the C tile (one
//                      read and one write of A22 per outer block).
// Look-ahead: the update of outer block K is split into the L-shaped part next to the diagonal
// (what chain K+1 needs) and the rest; chain K+1 runs on a second stream concurrently with the rest.
// All products use mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4: tcgen05 has no FP64 kind, so
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

// ---- (1) diagonal block: LU without exchanges + inverses of L (unit lower) and U.
// Thread i owns row i in REGISTERS (the j loops are fully unrolled so every register index is
// static); pivot rows are broadcast through shared memory, one barrier per step.
//   pass 1 (threads 0..63)   : right-looking elimination  A = L U, packed into shared memory;
//   pass 2 (threads 0..63)   : X = L^-1, rows from the top down      } concurrently, each pair of
//   pass 3 (threads 64..127) : Y = U^-1, rows from the bottom up     } warps on its own named barrier
constexpr int DS = NB + 1;   // row stride of the shared arrays (column reads are conflict free)
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__global__ void __launch_bounds__(2 * NB)
diag_kernel(double *__restrict__ A, int np, int k, double *__restrict__ dinv) {
  extern __shared__ __align__(16) double dsm[];
  double *LUs = dsm;                // packed LU: U on and above the diagonal, L below
  double *Xs = dsm + NB * DS;       // rows of L^-1
  double *Ys = dsm + 2 * NB * DS;   // rows of U^-1
  const int tid = threadIdx.x;
  double *Akk = A + ((size_t) k * NB) * np + (size_t) k * NB;
  double a[NB];
  if (tid < NB) {
    const double2 *row = reinterpret_cast<const double2 *>(Akk + (size_t) tid * np);
#pragma unroll
    for (int c = 0; c < NB / 2; c++) {
      const double2 v = row[c];
      a[2 * c] = v.x;
      a[2 * c + 1] = v.y;
    }
    // ---- pass 1
#pragma unroll
    for (int j = 0; j < NB; j++) {
      if (tid == j) {
#pragma unroll
        for (int c = j; c < NB; c++) LUs[j * DS + c] = a[c];
      }
      bar_named(1, NB);
      if (tid > j) {
        const double l = a[j] / LUs[j * DS + j];
        LUs[tid * DS + j] = l;
#pragma unroll
        for (int c = j + 1; c < NB; c++) a[c] = fma(-l, LUs[j * DS + c], a[c]);
      }
    }
  }
  __syncthreads();
  double *Li = dinv + (size_t) k * 2 * NB * NB, *Ui = Li + NB * NB;
#pragma unroll
  for (int c = 0; c < NB; c++) a[c] = 0.0;
  if (tid < NB) {
    // ---- pass 2: x = row tid of L^-1 (entries c < tid; the diagonal is 1)
#pragma unroll
    for (int j = 0; j < NB; j++) {
      if (tid == j) {
#pragma unroll
        for (int c = 0; c < j; c++) Xs[j * DS + c] = a[c];
        Xs[j * DS + j] = 1.0;
      }
      bar_named(1, NB);
      if (tid > j) {
        const double l = LUs[tid * DS + j];
#pragma unroll
        for (int c = 0; c <= j; c++) a[c] = fma(-l, Xs[j * DS + c], a[c]);
      }
    }
  } else {
    // ---- pass 3: y = row t of U^-1 (entries c >= t), rows from the bottom up
    const int t = tid - NB;
#pragma unroll
    for (int j = NB - 1; j >= 0; j--) {
      if (t == j) {
        const double inv = 1.0 / LUs[j * DS + j];
        Ys[j * DS + j] = inv;            // y[j] = 1 at its turn
#pragma unroll
        for (int c = j + 1; c < NB; c++) Ys[j * DS + c] = a[c] * inv;
      }
      bar_named(2, NB);
      if (t < j) {
        const double u = LUs[t * DS + j];
#pragma unroll
        for (int c = j; c < NB; c++) a[c] = fma(-u, Ys[j * DS + c], a[c]);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < NB * NB; e += 2 * NB) {
    const int r = e / NB, c = e % NB;
    Akk[(size_t) r * np + c] = LUs[r * DS + c];
    Li[e] = (c <= r) ? Xs[r * DS + c] : 0.0;
    Ui[e] = (c >= r) ? Ys[r * DS + c] : 0.0;
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

// ---- rhs: b[r] -= L21[r][k-panel] . y_k for every row below panel k (one warp per row)
__global__ void __launch_bounds__(256)
rhs_kernel(const double *__restrict__ A, int np, int k, double *__restrict__ b) {
  const int row = (k + 1) * NB + blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= np) return;
  const double *Ar = A + (size_t) row * np + (size_t) k * NB;
  const double *yk = b + (size_t) k * NB;
  double s = Ar[lane] * yk[lane] + Ar[lane + 32] * yk[lane + 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) b[row] -= s;
}

// ---- strip: after panel k (even), the 64x64 tiles of block column k+1 (rows > k) and of block row
// k+1 (columns > k+1) get  C -= L21 * U12  so that panel k+1 can be factored
__global__ void __launch_bounds__(128)
strip_kernel(double *__restrict__ A, int np, int k) {
  extern __shared__ __align__(16) double sm[];
  double *Ps = sm, *Qs = sm + NB * SP;
  const int nb = np / NB;
  const int ncol = nb - k - 1;                 // tiles (i, k+1), i = k+1 .. nb-1
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ti, tj;
  if ((int) blockIdx.x < ncol) { ti = k + 1 + blockIdx.x; tj = k + 1; }
  else { ti = k + 1; tj = k + 2 + (blockIdx.x - ncol); }
  const double *L = A + ((size_t) ti * NB) * np + (size_t) k * NB;    // L21 tile (ti, k)
  const double *U = A + ((size_t) k * NB) * np + (size_t) tj * NB;    // U12 tile (k, tj)
  double *Ct = A + ((size_t) ti * NB) * np + (size_t) tj * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int r = e / NB, c = e % NB;
    Ps[r * SP + c] = L[(size_t) r * np + c];
    Qs[r * SQ + c] = U[(size_t) r * np + c];
  }
  __syncthreads();
  double acc[4][4][2];
  tile_product_64(Ps, Qs, acc, warp, lane);
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32, g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      double2 *o = reinterpret_cast<double2 *>(Ct + (size_t) (wm + 8 * mt + g) * np + wn + 8 * nt + 2 * tq);
      double2 v = *o;
      v.x -= acc[mt][nt][0];
      v.y -= acc[mt][nt][1];
      *o = v;
    }
}

// ---- trailing update of outer block K:  C[128x128 tile] -= L[128x128] * U[128x128]
constexpr int TM = 128, TN = 128, TK = 128;
constexpr int KC = 16;                 // k-chunk per pipeline stage
constexpr int STAGES = 4;
constexpr int SA = KC + 4;             // 20: (row*20 + col) mod 16 distinct for row, col < 4  (conflict-free LDS.64)
constexpr int SB = TN + 4;             // 132: (k*132 + n) mod 16 distinct for k, n < 4
constexpr int STAGE_DOUBLES = TM * SA + KC * SB;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// mode 0: the L-shaped set of tiles next to the diagonal (block column K+1 from the diagonal down,
//         block row K+1 right of the diagonal) -- what the next chain needs;
// mode 1: the square (i, j >= K+2).   Tile indices in units of 128.
__global__ void __launch_bounds__(256)
update_kernel(double *__restrict__ A, int np, int K, int mode) {
  extern __shared__ __align__(16) double sm[];
  const int nB = np / TM;
  int ti, tj;
  if (mode == 0) {
    const int ncol = nB - K - 1;
    if ((int) blockIdx.x < ncol) { ti = K + 1 + blockIdx.x; tj = K + 1; }
    else { ti = K + 1; tj = K + 2 + (blockIdx.x - ncol); }
  } else {
    const int m = nB - K - 2;
    ti = K + 2 + blockIdx.x / m;
    tj = K + 2 + blockIdx.x % m;
  }
  const int row0 = ti * TM, col0 = tj * TN, kc = K * TK;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double *Lg = A + (size_t) row0 * np + kc;      // [128 rows][128 k]
  const double *Ug = A + (size_t) kc * np + col0;      // [128 k][128 cols]

  auto issue = [&](int chunk) {
    double *As = sm + (chunk % STAGES) * STAGE_DOUBLES;
    double *Bs = As + TM * SA;
    const int k0 = chunk * KC;
#pragma unroll
    for (int i = 0; i < 4; i++) {       // A chunk: 128 rows x 16 doubles = 1024 x 16 B
      const int r = (tid >> 3) + 32 * i, p = tid & 7;
      cp_async16(As + r * SA + 2 * p, Lg + (size_t) r * np + k0 + 2 * p);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {       // B chunk: 16 rows x 128 doubles = 1024 x 16 B
      const int kk = (tid >> 6) + 4 * i, p = tid & 63;
      cp_async16(Bs + kk * SB + 2 * p, Ug + (size_t) (k0 + kk) * np + 2 * p);
    }
  };
  constexpr int NCHUNK = TK / KC;
#pragma unroll
  for (int c = 0; c < STAGES - 1; c++) { issue(c); cp_async_commit(); }

  // 8 warps: 4 (m) x 2 (n); warp tile 32 x 64 = 4 x 8 DMMA tiles; accumulators start from C
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
  const int g = lane >> 2, tq = lane & 3;
  double acc[4][8][2];
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const double2 v = *reinterpret_cast<const double2 *>(A + (size_t) (row0 + wm + 8 * mt + g) * np + col0 + wn + 8 * nt + 2 * tq);
      acc[mt][nt][0] = v.x;
      acc[mt][nt][1] = v.y;
    }

  for (int c = 0; c < NCHUNK; c++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (c + STAGES - 1 < NCHUNK) issue(c + STAGES - 1);
    cp_async_commit();
    const double *As = sm + (c % STAGES) * STAGE_DOUBLES;
    const double *Bs = As + TM * SA;
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += 4) {
      double af[4], bf[8];
#pragma unroll
      for (int mt = 0; mt < 4; mt++) af[mt] = -As[(wm + 8 * mt + g) * SA + k0 + tq];
#pragma unroll
      for (int nt = 0; nt < 8; nt++) bf[nt] = Bs[(k0 + tq) * SB + wn + 8 * nt + g];
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
      *reinterpret_cast<double2 *>(A + (size_t) (row0 + wm + 8 * mt + g) * np + col0 + wn + 8 * nt + 2 * tq) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
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
  const size_t diag_smem = (size_t) 3 * NB * DS * sizeof(double);
  B200RT_CUDA(c, cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) diag_smem));
  const size_t update_smem = (size_t) STAGES * STAGE_DOUBLES * sizeof(double);
  B200RT_CUDA(c, cudaFuncSetAttribute(panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) panel_smem));
  B200RT_CUDA(c, cudaFuncSetAttribute(strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) panel_smem));
  B200RT_CUDA(c, cudaFuncSetAttribute(update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) update_smem));

  // second stream + events for the look-ahead
  if (!c->stream2) B200RT_CUDA(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  const int nK = np / TM;
  while ((int) c->lu_events.size() < 2 * nK + 1) {
    cudaEvent_t ev;
    B200RT_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->lu_events.push_back(ev);
  }
  cudaStream_t sB = c->stream2;
  auto chain = [&](int KB, cudaStream_t s) {       // factor panels 2KB and 2KB+1
    for (int h = 0; h < 2; h++) {
      const int k = 2 * KB + h;
      const int nrest = nblk - k - 1;
      diag_kernel<<<1, 2 * NB, diag_smem, s>>>(A, np, k, dinv);
      panel_kernel<<<2 * nrest + 1, 128, panel_smem, s>>>(A, np, k, dinv, b);
      launches += 2;
      if (nrest > 0) {
        rhs_kernel<<<(nrest * NB + 7) / 8, 256, 0, s>>>(A, np, k, b);
        launches++;
      }
      if (h == 0) {
        strip_kernel<<<2 * nrest - 1, 128, panel_smem, s>>>(A, np, k);   // nrest >= 1 here (np % 128 == 0)
        launches++;
      }
    }
  };
  cudaEvent_t ev_start = c->lu_events[2 * nK];
  B200RT_CUDA(c, cudaEventRecord(ev_start, st));
  B200RT_CUDA(c, cudaStreamWaitEvent(sB, ev_start, 0));
  chain(0, sB);
  B200RT_CUDA(c, cudaEventRecord(c->lu_events[0], sB));            // evP[0]
  for (int KB = 0; KB < nK; KB++) {
    B200RT_CUDA(c, cudaStreamWaitEvent(st, c->lu_events[2 * KB], 0)); // panels of KB factored
    const int m1 = nK - KB - 1;                                       // outer blocks after K
    if (m1 > 0) {
      update_kernel<<<2 * m1 - 1, 256, update_smem, st>>>(A, np, KB, 0);
      launches++;
      B200RT_CUDA(c, cudaEventRecord(c->lu_events[2 * KB + 1], st));  // evU[KB]
      B200RT_CUDA(c, cudaStreamWaitEvent(sB, c->lu_events[2 * KB + 1], 0));
      chain(KB + 1, sB);
      B200RT_CUDA(c, cudaEventRecord(c->lu_events[2 * KB + 2], sB));  // evP[KB+1]
      const int m2 = m1 - 1;
      if (m2 > 0) {
        update_kernel<<<m2 * m2, 256, update_smem, st>>>(A, np, KB, 1);
        launches++;
      }
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
