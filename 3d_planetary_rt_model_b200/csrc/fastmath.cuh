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

// exp(x) for finite x <= 0 (also correct up to x ~ +700), relative error <= 1e-12.  The power of two is added to the
// exponent field of the polynomial value (in [0.70, 1.42]); k is floored at -1000, so arguments below -693 return a
// value <= 2^-1000 ~ 1e-301 instead of walking through the denormals to 0 -- far below anything the marches resolve.
// Valid for |x| < 1.4e9 (the rint trick keeps k in int32): optical depths of the spherical grid are far below that.
__device__ __forceinline__ double exp_nonpos(double x) {
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
