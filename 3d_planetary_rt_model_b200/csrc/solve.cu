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
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>
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

// ---- (1) diagonal block: in-place BLOCK Gauss-Jordan inverse, 4 pivots per step, no pivoting.
// With J the 4 pivot indices of a step, P = A[J,J] and R the other 124 indices, one step is
//     A[R,R] -= A[R,J] P^-1 A[J,R];   A[J,R] <- P^-1 A[J,R];   A[R,J] <- -A[R,J] P^-1;   A[J,J] <- P^-1
// (the scalar in-place Gauss-Jordan step with a 4x4 pivot).  The rank-4 update of the whole block is exactly
// m8n8k4-shaped, so it runs on the FP64 tensor pipe: 512 threads = 16 warps in a 4 x 4 arrangement, each warp keeps
// a 32 x 32 piece of the block in DMMA accumulator layout (16 tiles, 32 registers per thread) for all 32 steps.
// Per step: the owners publish rows J and columns J through shared memory (double-buffered: ONE barrier per step,
// 32 barriers per inverse where the scalar version needed 128), every thread inverts the 4x4 pivot itself (cheaper
// than a second barrier), forms its B fragments (P^-1 A[J,R])[k = tq][column g] with 4 FMAs per tile column, issues 16
// DMMAs, and the owners overwrite the pivot rows / columns.  Measured: 67 us (scalar, 8x8 register tiles, DFMA) ->
// see profiles/ for the current figure; the inverse is the critical path of the last ~27 block columns.
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
__device__ __forceinline__ double sel4(double a0, double a1, double a2, double a3, int i) {
  const double lo = (i & 1) ? a1 : a0, hi = (i & 1) ? a3 : a2;
  return (i & 2) ? hi : lo;
}
// in-place inverse of a 4x4 matrix with a safely non-zero diagonal at every step (principal block of an M-matrix)
__device__ __forceinline__ void inv4(double (&m)[4][4]) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const double p = rcp_fast(m[j][j]);
#pragma unroll
    for (int c = 0; c < 4; c++) m[j][c] = (c == j) ? p : m[j][c] * p;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (i == j) continue;
      const double f = m[i][j];
#pragma unroll
      for (int c = 0; c < 4; c++) m[i][c] = (c == j) ? -f * p : fma(-f, m[j][c], m[i][c]);
    }
  }
}
constexpr int TB = 128;   // block size of the factorisation
constexpr int GJ_THREADS = 512;
__global__ void __launch_bounds__(GJ_THREADS)
gj128_kernel(const double *__restrict__ A, int np, int KB, double *__restrict__ dinv) {
  KB += blockIdx.x;     // (the factorisation launches one CTA; the profiling probe inverts many diagonal blocks at once)
  __shared__ __align__(16) double prow[2][4][TB];   // rows J of the step
  __shared__ __align__(16) double pcol[2][TB][4];   // columns J of the step
  __shared__ __align__(16) double pinv[2][4][4];    // P^-1 of the step
  __shared__ __align__(16) double praw[4][4];       // the next pivot block on its way to its inverse (owner warp only)
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int wr = w >> 2, wc = w & 3, wm = wr * 32, wn = wc * 32;
  const double *Akk = A + ((size_t) KB * TB) * np + (size_t) KB * TB;
  double acc[4][4][2];   // element (wm + 8 mt + g, wn + 8 nt + 2 tq + c)
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      const double2 v = *reinterpret_cast<const double2 *>(Akk + (size_t) (wm + 8 * mt + g) * np + wn + 8 * nt + 2 * tq);
      acc[mt][nt][0] = v.x;
      acc[mt][nt][1] = v.y;
    }
  // what the step after the current one needs, written into buffer `nb`: the inverse of its pivot block (formed by the
  // one warp that holds it, while the other warps are still in their DMMAs) and its pivot rows / columns.
  // band = 32-wide band of the pivots, tJ = tile index inside the band, half = which 4 of the tile's 8 indices.
  // (tJ is a compile-time constant at every call once the step loop is unrolled; the `t == tJ` loops keep the register
  // indices static)
  auto invert_next_pivot = [&](int band, int tJ, int half, int nb) {
    if (wr == band && wc == band) {               // warp-uniform
      if ((g >> 2) == half && (tq >> 1) == half) {
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (t == tJ) *reinterpret_cast<double2 *>(&praw[g & 3][(2 * tq) & 3]) = make_double2(acc[t][t][0], acc[t][t][1]);
      }
      __syncwarp();
      double m[4][4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double2 a = *reinterpret_cast<const double2 *>(&praw[q][0]);
        const double2 b = *reinterpret_cast<const double2 *>(&praw[q][2]);
        m[q][0] = a.x; m[q][1] = a.y; m[q][2] = b.x; m[q][3] = b.y;
      }
      __syncwarp();
      inv4(m);
      if (lane < 4) {
        *reinterpret_cast<double2 *>(&pinv[nb][lane][0]) =
            make_double2(sel4(m[0][0], m[1][0], m[2][0], m[3][0], lane), sel4(m[0][1], m[1][1], m[2][1], m[3][1], lane));
        *reinterpret_cast<double2 *>(&pinv[nb][lane][2]) =
            make_double2(sel4(m[0][2], m[1][2], m[2][2], m[3][2], lane), sel4(m[0][3], m[1][3], m[2][3], m[3][3], lane));
      }
    }
  };
  auto publish = [&](int band, int tJ, int half, int nb) {
    if (wr == band && (g >> 2) == half) {
#pragma unroll
      for (int t = 0; t < 4; t++)
        if (t == tJ) {
#pragma unroll
          for (int nt = 0; nt < 4; nt++)
            *reinterpret_cast<double2 *>(&prow[nb][g & 3][wn + 8 * nt + 2 * tq]) = make_double2(acc[t][nt][0], acc[t][nt][1]);
        }
    }
    if (wc == band && (tq >> 1) == half) {
#pragma unroll
      for (int t = 0; t < 4; t++)
        if (t == tJ) {
#pragma unroll
          for (int mt = 0; mt < 4; mt++)
            *reinterpret_cast<double2 *>(&pcol[nb][wm + 8 * mt + g][(2 * tq) & 3]) = make_double2(acc[mt][t][0], acc[mt][t][1]);
        }
    }
  };
  invert_next_pivot(0, 0, 0, 0);
  publish(0, 0, 0, 0);
  __syncthreads();
  for (int kk = 0; kk < 4; kk++) {         // 32-wide band holding the pivots
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {       // step within the band: tile index ks >> 1, half of the tile ks & 1
      const int tJ = ks >> 1, half = ks & 1, buf = ks & 1;
      const int tJn = ((ks + 1) & 7) >> 1, halfn = (ks + 1) & 1, kkn = kk + (ks == 7 ? 1 : 0);   // the step after
      const bool own_rows = (wr == kk) && ((g >> 2) == half);
      const bool own_cols = (wc == kk) && ((tq >> 1) == half);
      // B fragments of the update: (P^-1 A[J,:])[tq][wn + 8 nt + g]
      double pk[4], spb[4], fa[4];
      {
        const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][tq][0]);
        const double2 b = *reinterpret_cast<const double2 *>(&pinv[buf][tq][2]);
        pk[0] = a.x; pk[1] = a.y; pk[2] = b.x; pk[3] = b.y;
      }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        const int col = wn + 8 * nt + g;
        spb[nt] = pk[0] * prow[buf][0][col];
        spb[nt] = fma(pk[1], prow[buf][1][col], spb[nt]);
        spb[nt] = fma(pk[2], prow[buf][2][col], spb[nt]);
        spb[nt] = fma(pk[3], prow[buf][3][col], spb[nt]);
      }
#pragma unroll
      for (int mt = 0; mt < 4; mt++) fa[mt] = -pcol[buf][wm + 8 * mt + g][tq];
      // DMMA order: the tile holding the next pivot block first (its owner inverts it while the other DMMAs drain), then
      // the rest of that tile row / column (what the next step's publication reads), then the other nine tiles.
      dmma8x8x4(acc[tJn][tJn][0], acc[tJn][tJn][1], fa[tJn], spb[tJn]);
      if (kkn < 4) invert_next_pivot(kkn, tJn, halfn, buf ^ 1);
#pragma unroll
      for (int t = 0; t < 4; t++)
        if (t != tJn) {
          dmma8x8x4(acc[tJn][t][0], acc[tJn][t][1], fa[tJn], spb[t]);
          dmma8x8x4(acc[t][tJn][0], acc[t][tJn][1], fa[t], spb[tJn]);
        }
#pragma unroll
      for (int mt = 0; mt < 4; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
          if (mt != tJn && nt != tJn) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], fa[mt], spb[nt]);
      // the pivot rows / columns are then overwritten with their new values
      if (own_rows) {      // A[J,:] <- P^-1 A[J,:]
        double pr[4];
        {
          const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][g & 3][0]);
          const double2 b = *reinterpret_cast<const double2 *>(&pinv[buf][g & 3][2]);
          pr[0] = a.x; pr[1] = a.y; pr[2] = b.x; pr[3] = b.y;
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          double c0 = 0, c1 = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double2 r = *reinterpret_cast<const double2 *>(&prow[buf][q][wn + 8 * nt + 2 * tq]);
            c0 = fma(pr[q], r.x, c0);
            c1 = fma(pr[q], r.y, c1);
          }
          acc[tJ][nt][0] = c0;
          acc[tJ][nt][1] = c1;
        }
      }
      if (own_cols) {      // A[:,J] <- -A[:,J] P^-1, and the pivot block itself <- P^-1
        const int cq = (2 * tq) & 3;   // 0 or 2: this lane's two columns of J
        double pc0[4], pc1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][q][cq]);
          pc0[q] = a.x;
          pc1[q] = a.y;
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
          const double2 f01 = *reinterpret_cast<const double2 *>(&pcol[buf][wm + 8 * mt + g][0]);
          const double2 f23 = *reinterpret_cast<const double2 *>(&pcol[buf][wm + 8 * mt + g][2]);
          acc[mt][tJ][0] = -fma(f23.y, pc0[3], fma(f23.x, pc0[2], fma(f01.y, pc0[1], f01.x * pc0[0])));
          acc[mt][tJ][1] = -fma(f23.y, pc1[3], fma(f23.x, pc1[2], fma(f01.y, pc1[1], f01.x * pc1[0])));
        }
        if (own_rows) {
          acc[tJ][tJ][0] = sel4(pc0[0], pc0[1], pc0[2], pc0[3], g & 3);
          acc[tJ][tJ][1] = sel4(pc1[0], pc1[1], pc1[2], pc1[3], g & 3);
        }
      }
      if (kkn < 4) publish(kkn, tJn, halfn, buf ^ 1);
      __syncthreads();
    }
  }
  double *out = dinv + (size_t) KB * TB * TB;
#pragma unroll
  for (int mt = 0; mt < 4; mt++)
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
      *reinterpret_cast<double2 *>(out + (size_t) (wm + 8 * mt + g) * TB + wn + 8 * nt + 2 * tq) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
}

// ---- (1b) the same block Gauss-Jordan inverse on a CLUSTER of four CTAs (four SMs): CTA c keeps rows 32c .. 32c+31 of
// the block (8 warps x 16 columns, 8 DMMA tiles per warp).  On one SM a step costs ~3300 cycles: 1230 of DMMA issue and
// ~1500 of shared-memory bandwidth moving the fragments (measured, tools/dev/README.md); four SMs cut both by four.
// Per step the CTA that owns the pivot rows writes the raw rows J of the next step, and the warp that owns the next
// pivot block its inverse, into the shared memory of ALL four CTAs (distributed shared memory), every CTA publishes its
// own part of the columns J locally, and one cluster barrier closes the step.
#ifndef GJ_CLUSTER
#define GJ_CLUSTER 4
#endif
constexpr int GJC_THREADS = 256;
template <int GJC_CTAS>
__global__ void __launch_bounds__(GJC_THREADS)
gj128_cluster_kernel(const double *__restrict__ A, int np, int KB, double *__restrict__ dinv) {
  constexpr int GJC_ROWS = TB / GJC_CTAS, MT = GJC_ROWS / 8;   // rows and m-tiles per CTA (4 CTAs: 32, 4; 8 CTAs: 16, 2)
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int) cluster.block_rank();
  __shared__ __align__(16) double prow[2][4][TB];          // raw rows J of the step (same content in every CTA)
  __shared__ __align__(16) double pcol[2][GJC_ROWS][4];    // raw columns J, this CTA's rows
  __shared__ __align__(16) double pinv[2][4][4];           // P^-1 of the step (same content in every CTA)
  __shared__ __align__(16) double praw[4][4];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  const int wn = w * 16, row0 = crank * GJC_ROWS;
  const double *Akk = A + ((size_t) KB * TB) * np + (size_t) KB * TB;
  double acc[MT][2][2];   // element (row0 + 8 mt + g, wn + 8 nt + 2 tq + c)
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++) {
      const double2 v = *reinterpret_cast<const double2 *>(Akk + (size_t) (row0 + 8 * mt + g) * np + wn + 8 * nt + 2 * tq);
      acc[mt][nt][0] = v.x;
      acc[mt][nt][1] = v.y;
    }
  double *prow_r[GJC_CTAS], *pinv_r[GJC_CTAS];
#pragma unroll
  for (int r = 0; r < GJC_CTAS; r++) {
    prow_r[r] = cluster.map_shared_rank(&prow[0][0][0], r);
    pinv_r[r] = cluster.map_shared_rank(&pinv[0][0][0], r);
  }
  // step (band, tJ, half): pivots 32 band + 8 tJ + 4 half .. +3; they sit in CTA `band`, m-tile tJ, lanes g >> 2 == half,
  // and in warp wJ = (32 band + 8 tJ) / 16, n-tile nJ = tJ & 1 of every CTA, lanes tq >> 1 == half
  auto invert_next_pivot = [&](int band, int tJ, int half, int nb) {
    const int wJ = (GJC_ROWS * band + 8 * tJ) / 16;
    if (crank == band && w == wJ) {               // one warp of one CTA
      if ((g >> 2) == half && (tq >> 1) == half) {
#pragma unroll
        for (int t = 0; t < MT; t++)
          if (t == tJ) *reinterpret_cast<double2 *>(&praw[g & 3][(2 * tq) & 3]) = make_double2(acc[t][t & 1][0], acc[t][t & 1][1]);
      }
      __syncwarp();
      double m[4][4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double2 a = *reinterpret_cast<const double2 *>(&praw[q][0]);
        const double2 b = *reinterpret_cast<const double2 *>(&praw[q][2]);
        m[q][0] = a.x; m[q][1] = a.y; m[q][2] = b.x; m[q][3] = b.y;
      }
      __syncwarp();
      inv4(m);
      if (lane < 4) {
        const double2 lo = make_double2(sel4(m[0][0], m[1][0], m[2][0], m[3][0], lane), sel4(m[0][1], m[1][1], m[2][1], m[3][1], lane));
        const double2 hi = make_double2(sel4(m[0][2], m[1][2], m[2][2], m[3][2], lane), sel4(m[0][3], m[1][3], m[2][3], m[3][3], lane));
#pragma unroll
        for (int r = 0; r < GJC_CTAS; r++) {
          *reinterpret_cast<double2 *>(pinv_r[r] + (nb * 4 + lane) * 4 + 0) = lo;
          *reinterpret_cast<double2 *>(pinv_r[r] + (nb * 4 + lane) * 4 + 2) = hi;
        }
      }
    }
  };
  auto publish = [&](int band, int tJ, int half, int nb) {
    if (crank == band && (g >> 2) == half) {      // the rows J: into every CTA
#pragma unroll
      for (int t = 0; t < MT; t++)
        if (t == tJ) {
#pragma unroll
          for (int nt = 0; nt < 2; nt++) {
            const double2 v = make_double2(acc[t][nt][0], acc[t][nt][1]);
#pragma unroll
            for (int r = 0; r < GJC_CTAS; r++)
              *reinterpret_cast<double2 *>(prow_r[r] + ((size_t) nb * 4 + (g & 3)) * TB + wn + 8 * nt + 2 * tq) = v;
          }
        }
    }
    const int wJ = (GJC_ROWS * band + 8 * tJ) / 16;
    if (w == wJ && (tq >> 1) == half) {           // this CTA's part of the columns J: local
#pragma unroll
      for (int t = 0; t < 2; t++)
        if (t == (tJ & 1)) {
#pragma unroll
          for (int mt = 0; mt < MT; mt++)
            *reinterpret_cast<double2 *>(&pcol[nb][8 * mt + g][(2 * tq) & 3]) = make_double2(acc[mt][t][0], acc[mt][t][1]);
        }
    }
  };
  invert_next_pivot(0, 0, 0, 0);
  publish(0, 0, 0, 0);
  cluster.sync();
  constexpr int SPB = 2 * MT;              // steps per band (per CTA of pivot rows)
  for (int kk = 0; kk < GJC_CTAS; kk++) {  // CTA holding the pivot rows
#pragma unroll
    for (int ks = 0; ks < SPB; ks++) {     // step within its rows: m-tile ks >> 1, half ks & 1
      const int tJ = ks >> 1, half = ks & 1, buf = ks & 1, nJ = tJ & 1, wJ = (GJC_ROWS * kk + 8 * tJ) / 16;
      const int tJn = ((ks + 1) % SPB) >> 1, halfn = (ks + 1) & 1, kkn = kk + (ks == SPB - 1 ? 1 : 0);   // the step after
      const bool own_rows = (crank == kk) && ((g >> 2) == half);
      const bool own_cols = (w == wJ) && ((tq >> 1) == half);
      // B fragments of the update: (P^-1 A[J,:])[tq][wn + 8 nt + g];  A fragments: -A[:,J][row][tq]
      double pk[4], spb[2], fa[MT];
      {
        const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][tq][0]);
        const double2 b = *reinterpret_cast<const double2 *>(&pinv[buf][tq][2]);
        pk[0] = a.x; pk[1] = a.y; pk[2] = b.x; pk[3] = b.y;
      }
#pragma unroll
      for (int nt = 0; nt < 2; nt++) {
        const int col = wn + 8 * nt + g;
        spb[nt] = pk[0] * prow[buf][0][col];
        spb[nt] = fma(pk[1], prow[buf][1][col], spb[nt]);
        spb[nt] = fma(pk[2], prow[buf][2][col], spb[nt]);
        spb[nt] = fma(pk[3], prow[buf][3][col], spb[nt]);
      }
#pragma unroll
      for (int mt = 0; mt < MT; mt++) fa[mt] = -pcol[buf][8 * mt + g][tq];
      // the tile holding the next pivot block first: its owner inverts it while the other DMMAs drain
      dmma8x8x4(acc[tJn][tJn & 1][0], acc[tJn][tJn & 1][1], fa[tJn], spb[tJn & 1]);
      if (kkn < GJC_CTAS) invert_next_pivot(kkn, tJn, halfn, buf ^ 1);
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
          if (!(mt == tJn && nt == (tJn & 1))) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], fa[mt], spb[nt]);
      // the pivot rows / columns are then overwritten with their new values
      if (own_rows) {      // A[J,:] <- P^-1 A[J,:]
        double pr[4];
        {
          const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][g & 3][0]);
          const double2 b = *reinterpret_cast<const double2 *>(&pinv[buf][g & 3][2]);
          pr[0] = a.x; pr[1] = a.y; pr[2] = b.x; pr[3] = b.y;
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          double c0 = 0, c1 = 0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const double2 r = *reinterpret_cast<const double2 *>(&prow[buf][q][wn + 8 * nt + 2 * tq]);
            c0 = fma(pr[q], r.x, c0);
            c1 = fma(pr[q], r.y, c1);
          }
          acc[tJ][nt][0] = c0;
          acc[tJ][nt][1] = c1;
        }
      }
      if (own_cols) {      // A[:,J] <- -A[:,J] P^-1, and the pivot block itself <- P^-1
        const int cq = (2 * tq) & 3;   // 0 or 2: this lane's two columns of J
        double pc0[4], pc1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const double2 a = *reinterpret_cast<const double2 *>(&pinv[buf][q][cq]);
          pc0[q] = a.x;
          pc1[q] = a.y;
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          const double2 f01 = *reinterpret_cast<const double2 *>(&pcol[buf][8 * mt + g][0]);
          const double2 f23 = *reinterpret_cast<const double2 *>(&pcol[buf][8 * mt + g][2]);
          acc[mt][nJ][0] = -fma(f23.y, pc0[3], fma(f23.x, pc0[2], fma(f01.y, pc0[1], f01.x * pc0[0])));
          acc[mt][nJ][1] = -fma(f23.y, pc1[3], fma(f23.x, pc1[2], fma(f01.y, pc1[1], f01.x * pc1[0])));
        }
        if (own_rows) {
          acc[tJ][nJ][0] = sel4(pc0[0], pc0[1], pc0[2], pc0[3], g & 3);
          acc[tJ][nJ][1] = sel4(pc1[0], pc1[1], pc1[2], pc1[3], g & 3);
        }
      }
      if (kkn < GJC_CTAS) publish(kkn, tJn, halfn, buf ^ 1);
      cluster.sync();
    }
  }
  double *out = dinv + (size_t) KB * TB * TB;
#pragma unroll
  for (int mt = 0; mt < MT; mt++)
#pragma unroll
    for (int nt = 0; nt < 2; nt++)
      *reinterpret_cast<double2 *>(out + (size_t) (row0 + 8 * mt + g) * TB + wn + 8 * nt + 2 * tq) =
          make_double2(acc[mt][nt][0], acc[mt][nt][1]);
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
// Pipeline shapes (overridable for experiments).  Measured on the n = 5841 solve of the bench (ms; first pair = bulk
// update, second = the chain's panel / L-shaped update kernels, (k-chunk, cp.async stages)):
//   (8,4)+(16,4) 6.35   (8,3)+(16,4) 5.98   (8,2) 6.36   (16,3) 6.34   (16,2) 6.21   (8,6) 6.70
//   (8,3)+(16,3) 5.90   (8,3)+(16,2) 5.88   (8,3)+(32,2) 5.97
// A shorter pipeline fill per tile beats deeper prefetch: the bulk update has two CTAs per SM covering each other's load
// latency, and the chain kernels are latency-bound tiles whose first DMMA waits for STAGES-1 chunks.
#ifndef KC_BULK_V
#define KC_BULK_V 8
#define STAGES_BULK_V 3
#endif
#ifndef KC_CHAIN_V
#define KC_CHAIN_V 16
#define STAGES_CHAIN_V 2
#endif
constexpr int KC_DEFAULT = KC_CHAIN_V;         // k-chunk per pipeline stage (must be a multiple of 16 for the quarter tiles)
constexpr int STAGES_DEFAULT = STAGES_CHAIN_V;
constexpr int KC_BULK = KC_BULK_V, STAGES_BULK = STAGES_BULK_V;
constexpr size_t gemm_smem_bytes(int kc, int stages, int tn) { return (size_t) stages * (TM * (kc + 4) + kc * (tn + 4)) * sizeof(double); }

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// MODE 0: update, block column KB+1 from the diagonal down -- what the next chain needs:  C -= L21 * A12.  (With
//         gridDim.x = 2 m1 - 1 it also covers block row KB+1 right of the diagonal; the factorisation launches m1
//         blocks: that row is only read by the updates of the NEXT block column, so it rides with the bulk update)
// MODE 1: update, the rest: block rows i >= KB+1, block columns j >= KB+2
// MODE 2: panel,  tile (KB+1+blockIdx.x, KB):  L21 = A21 * A_KK^-1 -> Lbuf[row][0..127] (the L panels live in
//         a double-buffered side array: the updates of block column KB read one half while chain KB+1 fills the other)
// NT = DMMA n-tiles per warp: 8 -> 128x128 output tile per CTA (bulk update), 4 -> 128x64 (the two
// kernels on the critical path run as twice as many CTAs with half the latency; blockIdx.y picks the half).
// KC = k-chunk per pipeline stage, STAGES = cp.async stages.  The bulk update runs <1, 4, 8, 4>: 128x64 tiles (64
// accumulator registers per thread) and 66 kB of shared memory per CTA, so TWO CTAs are resident per SM and one's
// C-tile prologue / epilogue overlaps the other's DMMA stream (a lone 128x128 CTA per SM left the tensor pipe idle
// ~40 % of each tile: 26.5 us per tile against 16.7 us of DMMA issue).
template <int MODE, int NT, int KC = KC_DEFAULT, int STAGES = STAGES_DEFAULT>
__global__ void __launch_bounds__(256, (MODE == 1) ? 2 : 1)
gemm128_kernel(double *__restrict__ A, int np, int KB, const double *__restrict__ dinv, double *__restrict__ Lbuf) {
  constexpr int SA = KC + 4;             // 20 / 12: (row*SA + col) mod 16 distinct for row, col < 4  (conflict-free LDS.64)
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
  } else if (MODE == 1) {      // rows from KB+1 (the block row next to the diagonal is not on the critical path), columns from KB+2
    const int m = nB - KB - 2;
    ti = KB + 1 + blockIdx.x / m;
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
    constexpr int PA = KC / 2;          // 16-byte pieces per A row: the chunk is 128 rows x KC doubles
#pragma unroll
    for (int i = 0; i < TM * PA / 256; i++) {
      const int e = tid + 256 * i, r = e / PA, p = e % PA;
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

// ---- the whole back substitution in ONE launch: CTA I owns y_I; for K = nK-1 .. 0 CTA K forms x_K = A_KK^-1 y_K, publishes
// it and raises flag[K]; every CTA I < K waits for the flag and applies y_I -= A_IK x_K.  Critical path per block column: one
// flag round trip + two 128x128 mat-vecs, against a kernel launch + the same two mat-vecs before (46 launches, ~0.5 ms).
__global__ void __launch_bounds__(512)
backsolve_fused_kernel(const double *__restrict__ A, int np, int nK, const double *__restrict__ dinv, const double *__restrict__ b,
                       double *__restrict__ x, int *flags) {
  __shared__ double yk[TB], xk[TB];
  // CTAs are handed out in blockIdx order: the producers (large K) come first, so a waiting CTA's producers are
  // always resident or done
  const int I = nK - 1 - (int) blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < TB) yk[tid] = b[(size_t) I * TB + tid];
  extern __shared__ __align__(16) double sdinv[];       // A_II^-1, 128 kB: needed at the very end, fetched at the very start
  {
    const double2 *src = reinterpret_cast<const double2 *>(dinv + (size_t) I * TB * TB);
    double2 *dst = reinterpret_cast<double2 *>(sdinv);
    for (int i = tid; i < TB * TB / 2; i += blockDim.x) dst[i] = __ldcg(src + i);
  }
  __syncthreads();
  // 16 warps x 8 rows each.  The matrix tile of the NEXT product is fetched into registers BEFORE the wait for the
  // vector it multiplies (the factorisation is complete: every tile is final), so the critical path per block column is
  // flag -> 1 kB vector -> FMAs + shuffle reduction, not flag -> 128 kB tile.
  double m[8][TB / 32];
  auto fetch = [&](const double *M, size_t ld) {
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int q = 0; q < TB / 32; q++) m[i][q] = __ldcg(M + (size_t) (warp * 8 + i) * ld + lane + 32 * q);
  };
  auto apply = [&](const double *v, double (&out)[8]) {
    double vv[TB / 32];
#pragma unroll
    for (int q = 0; q < TB / 32; q++) vv[q] = v[lane + 32 * q];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      double sacc = 0;
#pragma unroll
      for (int q = 0; q < TB / 32; q++) sacc += m[i][q] * vv[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      out[i] = sacc;
    }
  };
  for (int K = nK - 1; K > I; K--) {
    fetch(A + ((size_t) I * TB) * np + (size_t) K * TB, np);
    if (tid == 0) {
      while (*reinterpret_cast<volatile int *>(flags + K) == 0) { }
      __threadfence();
    }
    __syncthreads();
    if (tid < TB) xk[tid] = __ldcg(x + (size_t) K * TB + tid);
    __syncthreads();
    double r[8];
    apply(xk, r);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 8; i++) yk[warp * 8 + i] -= r[i];
    }
    __syncthreads();
  }
  // x_I = A_II^-1 y_I from the copy of the block inverse staged in shared memory at the start
  double r[8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int q = 0; q < TB / 32; q++) m[i][q] = sdinv[(size_t) (warp * 8 + i) * TB + lane + 32 * q];
  apply(yk, r);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[(size_t) I * TB + warp * 8 + i] = r[i];
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    *reinterpret_cast<volatile int *>(flags + I) = 1;
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
  B200RT_CUDA(c, c->lu_flag.ensure((size_t) nK * sizeof(int)));
  int *flags = c->lu_flag.as<int>();
  auto Lbuf = [&](int KB) { return Lbuf0 + (size_t) (KB & 1) * np * TB; };
  int launches = 0;

  prepare_kernel<<<np, 256, 0, st>>>(K, n, np, branching, S0, A, b, margin);
  launches++;
  B200RT_CUDA(c, cudaGetLastError());
  // read-backs go through page-locked scratch (common.hpp, PinnedBuf): hm = margins [np], then hy / hr / hs0 [n] each
  B200RT_CUDA(c, c->host_scratch.ensure(4 * (size_t) np * sizeof(double)));
  double *hm = c->host_scratch.as<double>(), *hy = hm + np, *hr = hy + np, *hs0 = hr + np;
  B200RT_CUDA(c, cudaMemcpyAsync(hm, margin, np * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaStreamSynchronize(st));
  // NaN-safe: std::min / std::max drop a NaN operand, so a non-finite margin is tested for explicitly (a NaN row must
  // fail the dominance test AND the certificate, never slip through them)
  double min_margin = 1e300;
  bool finite_in = true;
  for (int i = 0; i < n; i++) {
    if (!std::isfinite(hm[i])) { finite_in = false; min_margin = -INFINITY; break; }
    min_margin = std::min(min_margin, hm[i]);
  }
  if (res) res->min_margin = min_margin;
  if (!finite_in)
    return fail(c, B200RT_ERR_NOT_DOMINANT, "I - w*K has non-finite entries (NaN / Inf in the influence matrix or the branching ratio)");
  if (!(min_margin > 0.0)) {
    // Not row dominant as it stands (the multiplet emissions: K[(v0,iu),(v,ju)] carries the ORIGIN voxel's lower-state
    // density, multiplet_CFR_emission.hpp:264-271, so its rows are probabilities only after a diagonal rescaling).
    // Certificate instead: K >= 0 and max_i (w K)^m 1 < 1 for some m  =>  rho(w K) < 1  =>  I - w K is a nonsingular
    // M-matrix: every Schur complement is again one, all pivots are positive, and elimination without row exchanges
    // is componentwise backward stable (its error bound is invariant under the diagonal rescaling that makes the
    // matrix row dominant).  x / rabs / margin are free at this point and serve as scratch.
    bool certified = false;
    double kmin = 0, ymax = 1e300;
    for (int i = 0; i < n; i++) hy[i] = 1.0;
    B200RT_CUDA(c, cudaMemcpyAsync(x, hy, n * sizeof(double), cudaMemcpyHostToDevice, st));
    B200RT_CUDA(c, cudaStreamSynchronize(st));     // hy is read back into below
    double *ya = x, *yb = rabs;
    const int max_iter = 2048;
    for (int it = 0; it < max_iter && !certified; it++) {
      kmatvec_kernel<<<(n + 7) / 8, 256, 0, st>>>(K, n, branching, ya, yb, it == 0 ? margin : nullptr);
      launches++;
      std::swap(ya, yb);
      if (it == 0) {
        B200RT_CUDA(c, cudaMemcpyAsync(hm, margin, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200RT_CUDA(c, cudaStreamSynchronize(st));
        for (int i = 0; i < n; i++) {
          if (!std::isfinite(hm[i])) { kmin = -INFINITY; break; }
          kmin = std::min(kmin, hm[i]);
        }
        if (!(kmin >= 0.0) || !(branching >= 0.0)) break;
      }
      if ((it & 7) == 7 || it == 0) {
        B200RT_CUDA(c, cudaMemcpyAsync(hy, ya, n * sizeof(double), cudaMemcpyDeviceToHost, st));
        B200RT_CUDA(c, cudaStreamSynchronize(st));
        ymax = 0;
        for (int i = 0; i < n; i++) {
          if (!std::isfinite(hy[i])) { ymax = INFINITY; break; }
          ymax = std::max(ymax, hy[i]);
        }
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

  const size_t gemm_smem = gemm_smem_bytes(KC_BULK, STAGES_BULK, TN / 2);             // 128x64 tiles, bulk update
  const size_t gemm_smem_h = gemm_smem_bytes(KC_DEFAULT, STAGES_DEFAULT, TN / 2);     // 128x64 tiles
  // quarter-width tiles (128x32) for the two kernels of the critical path once the trailing matrix is small: four
  // times as many CTAs, half the DMMA work each -- the tail of the factorisation is latency, not throughput
  const size_t gemm_smem_q = gemm_smem_bytes(KC_DEFAULT, STAGES_DEFAULT, TN / 4);
  if (!c->solve_attrs_set) {   // per device, not per call: a sweep solves thousands of small systems
    B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_h));
    B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<1, 4, KC_BULK, STAGES_BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem));
    B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_h));
    B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_q));
    B200RT_CUDA(c, cudaFuncSetAttribute(gemm128_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) gemm_smem_q));
    B200RT_CUDA(c, cudaFuncSetAttribute(backsolve_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (TB * TB * sizeof(double))));
    c->solve_attrs_set = true;
  }
  static const int quarter_below = getenv("B200RT_LU_QUARTER_BELOW") ? atoi(getenv("B200RT_LU_QUARTER_BELOW")) : 1024;
  // (quarter-width tiles for the chain kernels whenever fewer block columns than this remain.  Measured, n = 5841:
  //  never 6.15 ms, below 20: 5.94, below 36: 5.88, always: 5.80 -- with 58 kB of shared memory per CTA several quarter
  //  tiles share an SM, and 2 (2 m1 - 1) half tiles left a second, nearly empty wave)
  static const int beside_above = getenv("B200RT_LU_BESIDE_ABOVE") ? atoi(getenv("B200RT_LU_BESIDE_ABOVE")) : (1 << 30);

  // second stream + events for the look-ahead
  if (!c->stream2) {   // the chain is the critical path: give its stream the highest priority
    int lo = 0, hi = 0;
    B200RT_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200RT_CUDA(c, cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, hi));
  }
  while ((int) c->lu_events.size() < 3 * nK + 2) {
    cudaEvent_t ev;
    B200RT_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->lu_events.push_back(ev);
  }
  if (!c->stream3) {
    int lo = 0, hi = 0;
    B200RT_CUDA(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200RT_CUDA(c, cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, hi));
  }
  cudaStream_t sB = c->stream2, sU = c->stream3;
  const int launches_before = launches;
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
  // whether a 4-CTA cluster of the Gauss-Jordan kernel can be resident is decided ONCE, outside any stream capture (a
  // failed launch inside a capture would invalidate it); the one-SM kernel is the form for devices / partitions without
  static const bool use_cluster = [] {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(GJ_CLUSTER); cfg.blockDim = dim3(GJC_THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = GJ_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n_clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&n_clusters, gj128_cluster_kernel<GJ_CLUSTER>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return n_clusters > 0;
  }();
  auto chain = [&](int KB, cudaStream_t s) {       // invert the diagonal block, form L21
    const int m1 = nK - KB - 1;
    {   // diagonal-block inverse on a cluster of four SMs (gj128_cluster_kernel); gj128_kernel is the one-SM form
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(GJ_CLUSTER); cfg.blockDim = dim3(GJC_THREADS); cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = GJ_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      const double *Ac = A;
      if (use_cluster) cudaLaunchKernelEx(&cfg, gj128_cluster_kernel<GJ_CLUSTER>, Ac, np, KB, dinv);
      else gj128_kernel<<<1, GJ_THREADS, 0, s>>>(A, np, KB, dinv);
    }
    launches++;
    if (m1 > 0) {
      if (m1 <= quarter_below) gemm128_kernel<2, 2><<<dim3(m1, 4), 256, gemm_smem_q, s>>>(A, np, KB, dinv, Lbuf(KB));
      else gemm128_kernel<2, 4><<<dim3(m1, 2), 256, gemm_smem_h, s>>>(A, np, KB, dinv, Lbuf(KB));
      launches += 1;
    }
  };
  // The ~280 launches of the factorisation and back substitution (two streams, ~140 event dependencies) are captured
  // once per (np, workspace) into a CUDA graph and replayed: on the last ~27 block columns the critical path is a chain
  // of short kernels, and the graph removes the host launch gaps between them.
  auto record = [&]() -> cudaError_t {
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return e_; } while (0)
    // Three streams.  sB (high priority): the chain -- diagonal-block inverse + panel of block column KB+1.  sU (high
    // priority): the L-shaped update next to the diagonal that the chain waits for.  st: right-hand side + the bulk update.
    // The bulk update of block column KB needs only the panel of KB (evP[KB]), so it COULD run beside the L-shaped update
    // of KB (which, alone on the machine, costs 20-33 us of every bulk-bound step: tools/dev/solve_trace.py) -- measured,
    // it must not: with the bulk update's CTAs resident (two per SM, the whole register file) the L-shaped update and
    // the cluster launch of the next inverse wait for SMs to drain, and the solve goes 6.36 -> 7.13 ms, monotonically in
    // the number of block columns handled that way (B200RT_LU_BESIDE_ABOVE, tools/dev/README.md).  Default: never.
    // Events: evP[KB] = lu_events[3 KB] (panel of KB ready), evU[KB] = [3 KB + 1] (L-shaped update of KB done),
    // evB[KB] = [3 KB + 2] (bulk update + right-hand side of KB done).
    cudaEvent_t ev_start = c->lu_events[3 * nK];
    CK(cudaEventRecord(ev_start, st));
    CK(cudaStreamWaitEvent(sB, ev_start, 0));
    CK(cudaStreamWaitEvent(sU, ev_start, 0));
    mark("start", 0, st);
    mark("chain_begin", 0, sB);
    chain(0, sB);
    mark("chain_end", 0, sB);
    CK(cudaEventRecord(c->lu_events[0], sB));            // evP[0]
    for (int KB = 0; KB < nK; KB++) {
      const int m1 = nK - KB - 1;                                       // block columns after KB
      if (m1 > 0) {
        // L-shaped update: block column KB+1 from the diagonal down, block row KB+1 right of it.  Its tiles were last
        // written by the bulk update of KB-1; its L panel buffer is the one chain KB+1 will overwrite afterwards.
        CK(cudaStreamWaitEvent(sU, c->lu_events[3 * KB], 0));
        if (KB > 0) CK(cudaStreamWaitEvent(sU, c->lu_events[3 * (KB - 1) + 2], 0));
        mark("ui_begin", KB, sU);
        if (m1 <= quarter_below) gemm128_kernel<0, 2><<<dim3(m1, 4), 256, gemm_smem_q, sU>>>(A, np, KB, dinv, Lbuf(KB));
        else gemm128_kernel<0, 4><<<dim3(m1, 2), 256, gemm_smem_h, sU>>>(A, np, KB, dinv, Lbuf(KB));
        mark("ui_end", KB, sU);
        launches++;
        CK(cudaEventRecord(c->lu_events[3 * KB + 1], sU));  // evU[KB]
        CK(cudaStreamWaitEvent(sB, c->lu_events[3 * KB + 1], 0));
        mark("chain_begin", KB + 1, sB);
        chain(KB + 1, sB);
        mark("chain_end", KB + 1, sB);
        CK(cudaEventRecord(c->lu_events[3 * KB + 3], sB));  // evP[KB+1]
        // right-hand side and bulk update of KB: beside the two above while the bulk update is the longer leg; once the
        // trailing matrix is small the chain is the critical path and the bulk update queues behind the L-shaped update
        CK(cudaStreamWaitEvent(st, c->lu_events[3 * KB + (m1 > beside_above ? 0 : 1)], 0));
        rhs128_kernel<<<(m1 * TB + 7) / 8, 256, 0, st>>>(Lbuf(KB), np, KB, b);
        launches++;
        const int m2 = m1 - 1;
        if (m2 > 0) {
          gemm128_kernel<1, 4, KC_BULK, STAGES_BULK><<<dim3((m2 + 1) * m2, 2), 256, gemm_smem, st>>>(A, np, KB, dinv, Lbuf(KB));
          mark("uii_end", KB, st);
          launches++;
        }
        CK(cudaEventRecord(c->lu_events[3 * KB + 2], st));  // evB[KB]
      }
    }
    CK(cudaStreamWaitEvent(st, c->lu_events[3 * (nK - 1)], 0));   // the last diagonal block is inverted: join the chain
    CK(cudaEventRecord(c->lu_events[3 * nK + 1], sU));              // ... and the update stream (a capture must be rejoined)
    CK(cudaStreamWaitEvent(st, c->lu_events[3 * nK + 1], 0));
    mark("factor_end", 0, st);
    if (nK <= NUM_SMS) {     // every CTA resident at once: the fused back substitution
      CK(cudaMemsetAsync(flags, 0, nK * sizeof(int), st));
      backsolve_fused_kernel<<<nK, 512, TB * TB * sizeof(double), st>>>(A, np, nK, dinv, b, x, flags);
      launches++;
    } else {
      for (int KB = nK - 1; KB >= 0; KB--) {
        backsolve128_kernel<<<KB + 1, 512, 0, st>>>(A, np, KB, dinv, b, x);
        launches++;
      }
    }
    mark("backsolve_end", 0, st);
  return cudaGetLastError();
#undef CK
  };
  if (trace) {
    B200RT_CUDA(c, record());
  } else {
    if (c->lu_graph && (c->lu_graph_np != np || c->lu_graph_A != A || c->lu_graph_dinv != dinv)) {
      cudaGraphExecDestroy(c->lu_graph);
      c->lu_graph = nullptr;
    }
    if (!c->lu_graph) {
      cudaGraph_t graph = nullptr;
      launches = 0;
      B200RT_CUDA(c, cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const cudaError_t e_rec = record();
      const cudaError_t e_end = cudaStreamEndCapture(st, &graph);
      if (e_rec != cudaSuccess || e_end != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        B200RT_CUDA(c, e_rec != cudaSuccess ? e_rec : e_end);
      }
      const cudaError_t e_inst = cudaGraphInstantiate(&c->lu_graph, graph, 0);
      cudaGraphDestroy(graph);
      B200RT_CUDA(c, e_inst);
      c->lu_graph_np = np; c->lu_graph_A = A; c->lu_graph_dinv = dinv; c->lu_graph_launches = launches;
      launches = launches_before;
    }
    B200RT_CUDA(c, cudaGraphLaunch(c->lu_graph, st));
    launches += c->lu_graph_launches;
  }
  B200RT_CUDA(c, cudaGetLastError());
  B200RT_CUDA(c, cudaMemcpyAsync(S, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  residual_kernel<<<(n + 7) / 8, 256, 0, st>>>(K, n, branching, S0, S, rabs);
  launches++;
  B200RT_CUDA(c, cudaGetLastError());
  B200RT_CUDA(c, cudaMemcpyAsync(hr, rabs, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  B200RT_CUDA(c, cudaMemcpyAsync(hs0, S0, n * sizeof(double), cudaMemcpyDeviceToHost, st));
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
  bool finite_out = true;
  for (int i = 0; i < n; i++) {
    if (!std::isfinite(hr[i])) finite_out = false;
    rmax = std::max(rmax, hr[i]);
    smax = std::max(smax, std::fabs(hs0[i]));
  }
  const double rel = finite_out ? rmax / (smax > 0 ? smax : 1.0) : INFINITY;
  if (res) { res->residual = rel; res->launches = launches; }
  // a solution whose residual is not small is never reported as success (elimination without exchanges is backward
  // stable for the matrices the checks above admit: anything else means the input slipped past them)
  if (!(rel <= 1e-8))
    return fail(c, B200RT_ERR_NOT_DOMINANT, "solve: relative residual " + std::to_string(rel) + " of (I - w*K) S = S0 is not small");
  return B200RT_OK;
}

} // namespace b200rt
