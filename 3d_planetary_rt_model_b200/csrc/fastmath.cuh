// fastmath.cuh -- branch-free FP64 exp and division for the march kernels (sm_100a).
//
// Why not exp()/operator/ of the CUDA math library: in the march kernels they are ~all of the
// arithmetic, and ncu's opcode mix of the first version showed 23-26 % of the issued
// instructions were UMOV (the library's polynomial coefficients re-materialised through uniform
// registers), 10 % BRA/BSSY/BSYNC (slow-path guards of exp and of the IEEE division) and only
// ~35 % FP64.  Here the coefficients are constant-bank operands of the DFMAs and there is no
// slow path: arguments of the marches are finite and <= 0 (exp) and normal, positive (division).
// Accuracy: exp_nonpos 1e-12 relative (see below), div_pos <= 1 ulp, div_approx 1e-13, div_fast 1e-12.
// The parity bar of these quantities is 1e-6 relative; geometry never uses these.
#pragma once

namespace b200rt {
namespace fm {

// degree-8 minimax (Remez, relative error) polynomial of exp on [-ln2/2, ln2/2]: max relative error 7.8e-13.
// The marches multiply up to ~2000 such factors into a transmission, so the accumulated error stays below 2e-9
// against the 1e-6 parity bar; three DFMAs per exp cheaper than a full-precision (degree-11) polynomial.
__constant__ double EXPC[9] = {
    0x1.ffffffffff7a3p-1, 0x1.ffffffffd563bp-1, 0x1.0000000087f33p-1, 0x1.555555a081893p-3, 0x1.55555405128c7p-5,
    0x1.111082af6efa3p-7, 0x1.6c18b3381c643p-10, 0x1.a1a81babf5675p-13, 0x1.9eda958cc5598p-16};
__constant__ double EXPK[4] = {
    1.4426950408889634074,        // log2(e)
    6755399441055744.0,           // 1.5 * 2^52: adding it leaves rint(t) in the low mantissa bits
    -6.93147180369123816490e-01,  // -ln2 high part (low 21 mantissa bits zero: k*hi is exact)
    -1.90821492927058770002e-10}; // -ln2 low part


#ifndef B200RT_EXP_TABLE
#define B200RT_EXP_TABLE 0
#endif
#if B200RT_EXP_TABLE
// 2^(j/64), j = 0..63
__device__ const double EXPT[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
__constant__ double EXPK64[3] = {
    92.33248261689366,              // 64 * log2(e)
    -6.93147180369123816490e-01 / 64, // -ln2/64 high part (k * hi exact)
    -1.90821492927058770002e-10 / 64};
#endif
// exp(x) for finite x <= 0 (also correct up to x ~ +700), relative error <= 1e-12.  The power of two is added to the
// exponent field of the polynomial value (in [0.70, 1.42]); k is floored at -1000, so arguments below -693 return a
// value <= 2^-1000 ~ 1e-301 instead of walking through the denormals to 0 -- far below anything the marches resolve.
// Valid for |x| < 1.4e9 (the rint trick keeps k in int32): optical depths of the spherical grid are far below that.
__device__ __forceinline__ double exp_nonpos(double x) {
#if B200RT_EXP_TABLE
  // table variant: x = (64 e + j) ln2/64 + r, |r| <= ln2/128; exp(x) = 2^e * 2^(j/64) * (degree-4 Taylor of exp(r), 3.9e-14)
  const double t64 = fma(x, EXPK64[0], EXPK[1]);
  const int k64 = __double2loint(t64);
  const double kd64 = t64 - EXPK[1];
  double r64 = fma(kd64, EXPK64[1], x);
  r64 = fma(kd64, EXPK64[2], r64);
  const double tj = __ldg(&EXPT[k64 & 63]);
  const int e64 = max(k64 >> 6, -1000);
  double p64 = fma(r64, 1.0 / 24, 1.0 / 6);
  p64 = fma(p64, r64, 0.5);
  p64 = fma(p64, r64, 1.0);
  p64 = fma(p64, r64, 1.0);
  p64 *= tj;
  return __hiloint2double(__double2hiint(p64) + (e64 << 20), __double2loint(p64));
#endif
  const double t = fma(x, EXPK[0], EXPK[1]);
  int k = __double2loint(t);
  const double kd = t - EXPK[1];
  double r = fma(kd, EXPK[2], x);
  r = fma(kd, EXPK[3], r);
  double p = EXPC[8];
#pragma unroll
  for (int i = 7; i >= 0; i--) p = fma(p, r, EXPC[i]);
  k = max(k, -1000);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// the same for ANY x <= 0: x < -1000 (result 0 either way) is pinned to -1000 with integer compares on the high
// word.  Needed on the plane-parallel grid, where a near-horizontal ray has path lengths ~1e25 cm (tau ~ 1e16).
// Costs ~10 % more issue slots per exp, so the spherical-grid kernels use exp_nonpos.
__device__ __forceinline__ double exp_nonpos_guarded(double x) {
  if ((unsigned) __double2hiint(x) > 0xC08F4000u) x = -1000.0;
  return exp_nonpos(x);
}

// a / b for normal positive b (|b| in [1e-290, 1e290]), no special cases
__device__ __forceinline__ double div_pos(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  const double q = a * r;
  const double rem = fma(-b, q, a);
  return fma(rem, r, q);
}

// a / b to ~1e-13 relative (two Newton steps, no final correction): for quantities held to 1e-6
__device__ __forceinline__ double div_approx(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return a * r;
}

// a / b to ~1e-12 relative (rcp.approx keeps 20 mantissa bits; one Newton step squares the error): the per-wavelength
// (1 - exp(-tau)) / tau of the brightness march
__device__ __forceinline__ double div_fast(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  const double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return a * r;
}

// a / b with the reciprocal r ~ 1/b (<= 1 ulp) supplied: one correction step, result <= 1 ulp
__device__ __forceinline__ double div_by(double a, double b, double r) {
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// sqrt(a^2 + b^2 + c^2) for moderate magnitudes (no overflow / underflow guards), <= 1 ulp from hypot(hypot(a,b),c)
__device__ __forceinline__ double norm3(double a, double b, double c) { return sqrt(fma(a, a, fma(b, b, c * c))); }

} // namespace fm
} // namespace b200rt
