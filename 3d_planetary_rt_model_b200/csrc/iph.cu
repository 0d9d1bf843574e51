// iph.cu -- Quemerais interplanetary-hydrogen Lyman-alpha background on the device (sm_100a).
//
// COMPILE THIS FILE WITH -fmad=false AND WITHOUT --use_fast_math: the march takes its step counts
// from float comparisons (TOP's `IF (SAB.LE.NORME)`, the exit at ALT(KMAX)), so every product / sum is
// rounded separately, as on the host, to follow the same path as the CPU restatement.
//
// Restates, for the device (reference src/quemerais_IPH_model/):
//   quemerais_iph_model                                 iph_model_interface.cpp:19-82
//   BACKGROUND (constants, wind frame, loop over LOS)   ipbackgroundCFR_fun.f:1-324
//   INTENSM_PH / TOP / DEN / IPAL3M / T                 ipbackgroundCFR_fun.f:363-741
// The Fortran is REAL*4 throughout and serial, and re-reads the 195 kB table on every call
// (:107-164); here the table is parsed once per context and the lines of sight are independent
// threads.  Only what feeds the returned xsn(2) = FLN(2) (:318) is evaluated.
//
// Mapping: one thread per line of sight.  The three tables the value depends on (DANS, SO(:,:,2),
// SN(:,:,2): 3 x 59 x 19 floats = 13.5 kB) and the two axes sit in shared memory; the bracket
// searches are the Fortran's linear scans restated as lower-bound searches (same bracket).
// Bound: FP32 instruction throughput (~500 outer steps x ~21 inner density look-ups per LOS).
//
// Deviations from the Fortran (shared with oracle/iph_oracle.c, which documents them): ACOS through a
// fixed float polynomial with its argument clamped to [-1, 1]; indices clamped where the Fortran reads
// one element past a table with zero weight.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "common.hpp"

namespace b200rt {

namespace {

struct IphConst {
  int kmax, lmax;
  float ua, dpi, sig, dtap, gral, dinf_b, dinf_o;   // dinf of the step-length model (idb) and of the output model
  float x2, y2, z2;                                  // observer in the wind frame [m]
  float a11, a12, a21, a22, a23, a31, a32, a33;
};

__device__ __forceinline__ float asin_core(float a) {   // 0 <= a <= 0.5
  const float z = a * a;
  float p = 4.2163199048E-2f;
  p = p * z + 2.4181311049E-2f;
  p = p * z + 4.5470025998E-2f;
  p = p * z + 7.4953002686E-2f;
  p = p * z + 1.6666752422E-1f;
  return p * z * a + a;
}
// one evaluation path for every lane (the three ranges of the CPU restatement folded into selects; the float
// operations that produce the result are the same ones, so the value is bit-identical)
__device__ __forceinline__ float acos_poly(float x) {
  const float PIO2 = 1.5707963267948966f, PI_F = 3.14159265358979f;
  if (x > 1.f) x = 1.f;
  if (x < -1.f) x = -1.f;
  const float a = fabsf(x);
  const bool big = a > 0.5f;
  const float s = asin_core(big ? sqrtf(0.5f * (1.f - a)) : a);
  const float two_s = 2.f * s;
  const float r_big = (x > 0.f) ? two_s : PI_F - two_s;
  const float r_small = (x >= 0.f) ? PIO2 - s : PIO2 + s;
  return big ? r_big : r_small;
}

struct Tables {
  const float *alt, *ang, *dans, *so, *sn;   // shared memory
  const float *dma;                           // dma[k] = (alt[k] - alt[k-1]) / 3 / UA: TOP's step limit (:473-476)
  int kmax, lmax;
};
// Where the previous look-up of this line of sight landed.  Consecutive look-ups are a fraction of a table cell
// apart (TOP steps by a third of the altitude cell at most), so the lower-bound searches start from the previous
// bracket and walk: 0-1 steps instead of the 5-6 halvings of a binary search.  The bracket found is the same.
struct Hint { int alt, ang; };

// lower bound (first j with x <= ax[j]) starting from the previous bracket: the hint is right or off by one in all but
// a handful of look-ups per line of sight.  Right: the two table values read to find that out ARE the bracket.  Off by
// one: one more value decides whether the neighbouring cell is the bracket.  Anything else (the first look-up, the jump
// from the end of one TOP segment to the next outer point) takes the binary search.  below = ax[lo-1], at = ax[lo].
struct Bracket { int lo; float below, at; };
__device__ __forceinline__ Bracket lower_bound_hint(const float *ax, int n, float x, int hint) {
  int lo = hint;
  float at = ax[lo];
  float below = ax[max(lo - 1, 0)];
  const bool up = at < x;
  const bool down = (lo > 0) && !(below < x);
  if (up || down) {
    const int nlo = up ? lo + 1 : lo - 1;
    const float extra = ax[up ? min(lo + 1, n - 1) : max(lo - 2, 0)];
    const bool ok = up ? (nlo >= n || !(extra < x)) : (nlo == 0 || extra < x);
    if (ok) {
      if (up) { below = at; at = extra; } else { at = below; below = extra; }
      lo = nlo;
    } else {
      int l = 0, hh = n;
      while (l < hh) {
        const int mid = (l + hh) >> 1;
        if (ax[mid] < x) l = mid + 1; else hh = mid;
      }
      lo = l;
      at = ax[min(lo, n - 1)];
      below = ax[max(lo - 1, 0)];
    }
  }
  return {lo, below, at};
}
// first j with T <= ang[j]  (the Fortran's arithmetic-IF scan, :517-528) and the weight of the upper neighbour;
// the case analysis of the scan (below the axis, on a node, beyond the axis) as selects
__device__ __forceinline__ void bracket_ang(const Tables &t, float T, int &ll, int &llp, float &dt, Hint &h) {
  const Bracket b = lower_bound_hint(t.ang, t.lmax, T, h.ang);
  const int lo = b.lo;
  h.ang = min(lo, t.lmax - 1);
  const bool on_node = (lo < t.lmax) && (T == b.at);
  ll = min(max(lo - 1, 0), t.lmax - 2);
  llp = ll + 1;
  float a0 = b.below, a1 = b.at;
  if (lo == 0 || lo >= t.lmax) { a0 = t.ang[ll]; a1 = t.ang[llp]; }     // off the axis: the end cell extrapolates
  dt = (T - a0) / (a1 - a0);
  if (on_node) { ll = lo; llp = min(lo + 1, t.lmax - 1); dt = 0.f; }
}
__device__ __forceinline__ void bracket_alt(const Tables &t, float Z, int &kk, int &kkp, float &du, Hint &h) {
  const Bracket b = lower_bound_hint(t.alt, t.kmax, Z, h.alt);
  const int lo = b.lo;
  h.alt = min(lo, t.kmax - 1);
  const bool beyond = lo >= t.kmax;
  const bool on_node = !beyond && (Z == b.at);
  kk = min(max(lo - 1, 0), t.kmax - 2);
  kkp = kk + 1;
  float a0 = b.below, a1 = b.at;
  if (lo == 0 || beyond) { a0 = t.alt[kk]; a1 = t.alt[kkp]; }
  du = (Z - a0) / (a1 - a0);
  if (on_node) { kk = lo; kkp = min(lo + 1, t.kmax - 1); du = 0.f; }
  if (beyond) { kk = kkp = t.kmax - 1; du = 0.f; }
}

__device__ __forceinline__ float den(const Tables &t, float Z, float T, int &ko, Hint &h) {   // DEN :492-547
  ko = 1;
  if (Z < t.alt[0]) return 0.f;
  if (Z > t.alt[t.kmax - 1]) Z = t.alt[t.kmax - 1];
  int ll, llp, kk, kkp;
  float dt, du;
  bracket_ang(t, T, ll, llp, dt, h);
  bracket_alt(t, Z, kk, kkp, du, h);
  const float a = t.dans[kk * t.lmax + ll], b = t.dans[kkp * t.lmax + ll];
  const float c = t.dans[kk * t.lmax + llp], d = t.dans[kkp * t.lmax + llp];
  const float fl = a + du * (b - a);
  const float flp = c + du * (d - c);
  ko = kk + 1;
  return fl + dt * (flp - fl);
}

__device__ __forceinline__ float holstein_T(float TO) {   // T :363-397
  if (TO < 0.f) return 0.f;
  if (TO <= 2.f) {
    float TN = 1.f, DTN = 1.f, Q = 1.f;
    do {
      DTN = -DTN * TO / sqrtf(Q * (Q + 1.f));
      TN = TN + DTN;
      Q = Q + 1.f;
    } while (Q < 12.f);
    return TN;
  }
  const float DEPI = 2.f / sqrtf(3.14159265358f);
  const float DX = 0.4f;
  float T = 0.f;
  if (TO < 600.f) T = DEPI * expf(-TO) * 0.5f * DX;
  for (int k = 1; k <= 10; k++) {
    const float X = k * DX;
    const float XU = -X * X;
    const float U = expf(XU);
    const float UU = TO * U;
    float DT = 0.f;
    if (UU < 600.f) DT = DEPI * U * expf(-UU);
    T = T + DT * DX;
  }
  return T;
}

__device__ float top(const Tables &t, const IphConst &c, float XF, float YF, float ZF, float XH, float YH, float ZH, Hint &h) {   // TOP :399-490
  const float UA = c.ua;
  float XA = XF / UA, XB = XH / UA, YA = YF / UA, YB = YH / UA, ZA = ZF / UA, ZB = ZH / UA;
  const float altp = t.alt[0] / UA;
  float RA = sqrtf(XA * XA + YA * YA + ZA * ZA);
  float RB = sqrtf(XB * XB + YB * YB + ZB * ZB);
  if (RA <= altp && RB <= altp) return 0.f;
  if (RA > RB) {
    float d;
    d = XA; XA = XB; XB = d;
    d = YA; YA = YB; YB = d;
    d = ZA; ZA = ZB; ZB = d;
    d = RA; RA = RB; RB = d;
  }
  float XAB = XB - XA, YAB = YB - YA, ZAB = ZB - ZA;
  const float NORME = sqrtf(XAB * XAB + YAB * YAB + ZAB * ZAB);
  if (NORME < .01f) return 0.f;
  XAB = XAB / NORME; YAB = YAB / NORME; ZAB = ZAB / NORME;
  const float DSA0 = NORME / 20.f;
  float TA = acos_poly(YA / RA) / c.dpi;
  int KP;
  float DN1 = den(t, RA * UA, TA, KP, h);
  DN1 = c.dinf_b * DN1;
  if (KP == t.kmax) KP = KP - 1;
  float DMA = t.dma[KP];
  float DSAB = fminf(DMA, DSA0);
  float SAB = 0.f, DT = 0.f;
  do {
    XA = XA + DSAB * XAB;
    YA = YA + DSAB * YAB;
    ZA = ZA + DSAB * ZAB;
    SAB = SAB + DSAB;
    RA = sqrtf(XA * XA + YA * YA + ZA * ZA);
    TA = acos_poly(YA / RA) / c.dpi;
    float DN = den(t, RA * UA, TA, KP, h);
    DN = c.dinf_b * DN;
    DT = DT + (DN + DN1) * .5f * DSAB * c.sig * UA;
    DN1 = DN;
    if (KP == t.kmax) KP = KP - 1;
    DMA = t.dma[KP];
    DSAB = fminf(DMA, DSA0);
  } while (SAB <= NORME);
  return DT;
}

__global__ void __launch_bounds__(128)
iph_kernel(IphConst c, const float *__restrict__ g_alt, const float *__restrict__ g_ang,
           const float *__restrict__ g_dans, const float *__restrict__ g_so, const float *__restrict__ g_sn,
           int n_los, const float *__restrict__ u1, const float *__restrict__ v1, const float *__restrict__ w1,
           float *__restrict__ fln, int *__restrict__ n_steps) {
  extern __shared__ float smf[];
  const int nt = c.kmax * c.lmax;
  float *s_alt = smf, *s_ang = s_alt + c.kmax, *s_dans = s_ang + c.lmax, *s_so = s_dans + nt, *s_sn = s_so + nt;
  float *s_dma = s_sn + nt;
  for (int i = threadIdx.x; i < c.kmax; i += blockDim.x) {
    s_alt[i] = g_alt[i];
    s_dma[i] = (i > 0) ? (g_alt[i] - g_alt[i - 1]) / 3.f / c.ua : 0.f;
  }
  for (int i = threadIdx.x; i < c.lmax; i += blockDim.x) s_ang[i] = g_ang[i];
  for (int i = threadIdx.x; i < nt; i += blockDim.x) { s_dans[i] = g_dans[i]; s_so[i] = g_so[i]; s_sn[i] = g_sn[i]; }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_los) return;
  Tables t = {s_alt, s_ang, s_dans, s_so, s_sn, s_dma, c.kmax, c.lmax};
  Hint h = {c.kmax - 1, 0}, h_out = {c.kmax - 1, 0};   // inner march (TOP) and outer march keep their own brackets

  const float U = c.a11 * u1[i] + c.a12 * v1[i];                       // :297-299
  const float V = c.a21 * u1[i] + c.a22 * v1[i] + c.a23 * w1[i];
  const float W = c.a31 * u1[i] + c.a32 * v1[i] + c.a33 * w1[i];
  const float X = c.x2, Y = c.y2, Z = c.z2;

  // INTENSM_PH :550-657
  float S = 0.f, TT = 0.f, FLN = 0.f;
  int steps = 0;
  const float RR = sqrtf(X * X + Y * Y + Z * Z);
  if (RR <= t.alt[t.kmax - 1]) {
    float YP = Y, R = RR, XAV = X, YAV = Y, ZAV = Z;
    float FOO = 0.f;   // IPAL3M zeroes F and CT before its early return but not FOO (:685-690): inside the innermost node and
                       // at the outer edge the caller's FOO keeps the value of the previous step
    for (;;) {
      const float TETA = acos_poly(YP / R) / c.dpi;
      int KO;
      const float DNA = den(t, R, TETA, KO, h_out);
      float DN1 = c.dinf_b * DNA;
      if (DN1 == 0.f) DN1 = 1.f;
      float DP = c.dtap * 0.05f / DN1;
      float DUA;
      if (KO < t.kmax) DUA = (t.alt[KO] - t.alt[KO - 1]) / 2.f;
      else DUA = (t.alt[t.kmax - 1] - t.alt[t.kmax - 2]) / 2.f;
      DP = fminf(DP, DUA);
      DP = fmaxf(DP, c.ua / 10.f);
      S = S + DP;
      const float XP = X + S * U;
      YP = Y + S * V;
      const float ZP = Z + S * W;
      R = sqrtf(XP * XP + YP * YP + ZP * ZP);
      if (R > t.alt[t.kmax - 1]) break;
      const float TETA2 = acos_poly(YP / R) / c.dpi;
      // IPAL3M :659-741 for the output density index
      float FN = 0.f;
      if (!(R < t.alt[0] || R >= t.alt[t.kmax - 1])) {
        int ll, llp, kk, kkp;
        float dt, du;
        bracket_ang(t, TETA2, ll, llp, dt, h_out);
        bracket_alt(t, R, kk, kkp, du, h_out);
        float fl = t.sn[kk * t.lmax + ll] + du * (t.sn[kkp * t.lmax + ll] - t.sn[kk * t.lmax + ll]);
        float flp = t.sn[kk * t.lmax + llp] + du * (t.sn[kkp * t.lmax + llp] - t.sn[kk * t.lmax + llp]);
        FN = fl + dt * (flp - fl);
        fl = t.so[kk * t.lmax + ll] + du * (t.so[kkp * t.lmax + ll] - t.so[kk * t.lmax + ll]);
        flp = t.so[kk * t.lmax + llp] + du * (t.so[kkp * t.lmax + llp] - t.so[kk * t.lmax + llp]);
        FOO = fl + dt * (flp - fl);
      }
      const float DTT = top(t, c, XAV, YAV, ZAV, XP, YP, ZP, h);
      const float cosff = (U * XP + V * YP + W * ZP) / R;
      const float corec = 0.25f * cosff * cosff + (11.f / 12.f);
      TT = TT + DTT * c.dinf_o / c.dinf_b;
      const float FFNN = FN + FOO * (corec - 1.f);
      const float TTTII = holstein_T(TT);
      const float DFLNC = FFNN * c.gral * TTTII * DP;
      FLN = FLN + DFLNC;
      XAV = XP; YAV = YP; ZAV = ZP;
      steps++;
    }
  }
  fln[i] = FLN;
  if (n_steps) n_steps[i] = steps;
}

} // namespace

// ---- host: constants of BACKGROUND (:176-235), evaluated in float with the host libm
int iph_set_table(b200rt_ctx *c, int kmax, int lmax, int ninf, float temp, const float *alt_au, const float *ang,
                  const float *dans, const float *sot, const float *so, const float *sn, const float *dinf_cm3) {
  (void) sot;   // the optically thin source function only feeds xot, which the reference discards (:318)
  if (kmax < 2 || kmax > 59 || lmax < 2 || lmax > 19 || ninf < 2 || ninf > 5 || !alt_au || !ang || !dans || !so || !sn || !dinf_cm3)
    return fail(c, B200RT_ERR_ARG, "b200rt_iph_set_table: bad table dimensions or null array");
  IphTable &T = c->iph;
  T.kmax = kmax; T.lmax = lmax;
  T.ua = 1.4959E+11f;
  T.dinf_b = dinf_cm3[0] * 1.E6f;
  T.dinf_o = dinf_cm3[1] * 1.E6f;
  const float XLA = 1.21566E-05f, PTF = 0.4162f;
  const float AM = 1.67333E-27f, BOLK = 1.38046E-23f;
  const float PY = 4 * atanf(1.f);
  T.dpi = PY / 180.f;
  const float SPI = sqrtf(PY);
  const float XNUZ = 1.f / XLA;
  const float E2 = 23.0677E-20f, EMAS = 9.1084E-28f;
  const float C = 2.99793E+10f;
  const float DLDN = XLA * XLA * 1.E+08f / C;
  const float SIGMAN = PY * E2 * PTF / (EMAS * C);
  T.sigmaf = SIGMAN * DLDN;
  const float DELNUD = XNUZ * sqrtf(2.f * BOLK * temp / AM) * 1.E+2f;
  float SIG = SIGMAN / (SPI * DELNUD);
  SIG = SIG * 1.E-4f;
  T.sig = SIG;
  T.dtap = 1.f / SIG;
  const float ALAMVENT = 252.3f * T.dpi, DECVENT = 8.7f * T.dpi;
  T.a[0] = sinf(ALAMVENT);  T.a[1] = -cosf(ALAMVENT);  T.a[2] = 0.f;
  T.a[3] = cosf(DECVENT) * cosf(ALAMVENT);  T.a[4] = cosf(DECVENT) * sinf(ALAMVENT);  T.a[5] = sinf(DECVENT);
  T.a[6] = -sinf(DECVENT) * cosf(ALAMVENT); T.a[7] = -sinf(DECVENT) * sinf(ALAMVENT); T.a[8] = cosf(DECVENT);
  const int nt = kmax * lmax;
  std::vector<float> h(kmax + lmax + 3 * (size_t) nt);
  for (int k = 0; k < kmax; k++) h[k] = alt_au[k] * T.ua;
  for (int l = 0; l < lmax; l++) h[kmax + l] = ang[l];
  std::memcpy(h.data() + kmax + lmax, dans, nt * sizeof(float));
  std::memcpy(h.data() + kmax + lmax + nt, so + (size_t) 1 * nt, nt * sizeof(float));       // SO(:,:,2)
  std::memcpy(h.data() + kmax + lmax + 2 * (size_t) nt, sn + (size_t) 1 * nt, nt * sizeof(float));   // SN(:,:,2)
  B200RT_CUDA(c, T.dev.ensure(h.size() * sizeof(float)));
  B200RT_CUDA(c, cudaMemcpyAsync(T.dev.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  T.loaded = true;
  return B200RT_OK;
}

// the READ sequence of BACKGROUND (:107-164): list-directed, i.e. a stream of numeric tokens
int iph_load_table(b200rt_ctx *c, const char *fname) {
  FILE *f = fopen(fname, "r");
  if (!f) return fail(c, B200RT_ERR_ARG, std::string("cannot open IPH table ") + fname);
  std::string buf;
  char tmp[65536];
  size_t got;
  while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) buf.append(tmp, got);
  fclose(f);
  const char *p = buf.c_str();
  const char *end = p + buf.size();
  bool ok = true;
  auto next = [&]() -> float {
    char *q;
    const float v = strtof(p, &q);
    if (q == p || q > end) { ok = false; return 0.f; }
    p = q;
    return v;
  };
  const int kmax = (int) next(), lmax = (int) next(), ninf = (int) next();
  if (!ok || kmax < 2 || kmax > 59 || lmax != 19 || ninf < 2 || ninf > 5)
    return fail(c, B200RT_ERR_ARG, "IPH table: unexpected KMAX LMAX INF header");
  const int nt = kmax * lmax;
  std::vector<float> alt(kmax), ang(lmax), dans(nt), sot(nt), so((size_t) ninf * nt), sn((size_t) ninf * nt), dinf(ninf);
  float hdr[8], temp = 0;
  for (float &v : hdr) v = next();
  temp = hdr[3];
  dinf[0] = hdr[7];
  const int c0[4] = {0, 5, 10, 15}, c1[4] = {5, 10, 15, 19};
  float *first[4] = {dans.data(), sot.data(), so.data(), sn.data()};
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      for (int l = c0[b]; l < c1[b]; l++) ang[l] = next();
      for (int k = 0; k < kmax; k++) {
        alt[k] = next();
        for (int l = c0[b]; l < c1[b]; l++) first[a][k * lmax + l] = next();
      }
    }
  for (int ii = 1; ii < ninf; ii++) {
    for (float &v : hdr) v = next();
    temp = hdr[3];
    dinf[ii] = hdr[7];
    float *arr[2] = {so.data() + (size_t) ii * nt, sn.data() + (size_t) ii * nt};
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 4; b++) {
        for (int l = c0[b]; l < c1[b]; l++) ang[l] = next();
        for (int k = 0; k < kmax; k++) {
          (void) next();   // ZALT: overwritten by ALT*UA afterwards (:181)
          for (int l = c0[b]; l < c1[b]; l++) arr[a][k * lmax + l] = next();
        }
      }
  }
  if (!ok) return fail(c, B200RT_ERR_ARG, "IPH table: file ended before the READ sequence was complete");
  return iph_set_table(c, kmax, lmax, ninf, temp, alt.data(), ang.data(), dans.data(), sot.data(), so.data(), sn.data(),
                       dinf.data());
}

int iph_background(b200rt_ctx *c, float fs, float xpos, float ypos, float zpos, int n_los, const float *u1,
                   const float *v1, const float *w1, float *fln, int *n_steps) {
  IphTable &T = c->iph;
  if (!T.loaded) return fail(c, B200RT_ERR_STATE, "IPH table not loaded (b200rt_iph_load_table / b200rt_iph_set_table)");
  if (n_los <= 0 || !u1 || !v1 || !w1 || !fln) return fail(c, B200RT_ERR_ARG, "b200rt_iph_background: bad argument");
  IphConst k;
  k.kmax = T.kmax; k.lmax = T.lmax; k.ua = T.ua; k.dpi = T.dpi; k.sig = T.sig; k.dtap = T.dtap;
  const float GZERO = fs * T.sigmaf;
  k.gral = GZERO * 1.E-10f;
  k.dinf_b = T.dinf_b; k.dinf_o = T.dinf_o;
  k.a11 = T.a[0]; k.a12 = T.a[1]; k.a21 = T.a[3]; k.a22 = T.a[4]; k.a23 = T.a[5]; k.a31 = T.a[6]; k.a32 = T.a[7]; k.a33 = T.a[8];
  k.x2 = (k.a11 * xpos + k.a12 * ypos) * T.ua;                         // :270-275
  k.y2 = (k.a21 * xpos + k.a22 * ypos + k.a23 * zpos) * T.ua;
  k.z2 = (k.a31 * xpos + k.a32 * ypos + k.a33 * zpos) * T.ua;
  const size_t n = (size_t) n_los;
  B200RT_CUDA(c, T.io.ensure(n * (4 * sizeof(float) + sizeof(int))));
  float *d_u = T.io.as<float>(), *d_v = d_u + n, *d_w = d_v + n, *d_f = d_w + n;
  int *d_s = reinterpret_cast<int *>(d_f + n);
  B200RT_CUDA(c, cudaMemcpyAsync(d_u, u1, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  B200RT_CUDA(c, cudaMemcpyAsync(d_v, v1, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  B200RT_CUDA(c, cudaMemcpyAsync(d_w, w1, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  const int nt = T.kmax * T.lmax;
  const float *tab = T.dev.as<float>();
  const size_t smem = (size_t) (2 * T.kmax + T.lmax + 3 * nt) * sizeof(float);   // axes, step table, DANS, SO, SN
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, c->stream);
  iph_kernel<<<(n_los + 127) / 128, 128, smem, c->stream>>>(k, tab, tab + T.kmax, tab + T.kmax + T.lmax,
                                                           tab + T.kmax + T.lmax + nt, tab + T.kmax + T.lmax + 2 * (size_t) nt,
                                                           n_los, d_u, d_v, d_w, d_f, n_steps ? d_s : nullptr);
  cudaEventRecord(e1, c->stream);
  B200RT_CUDA(c, cudaGetLastError());
  B200RT_CUDA(c, cudaMemcpyAsync(fln, d_f, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (n_steps) B200RT_CUDA(c, cudaMemcpyAsync(n_steps, d_s, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  c->phase_ms[PH_IPH] = ms;
  c->phase_launches[PH_IPH] = 1;
  return B200RT_OK;
}

// quemerais_iph_model (iph_model_interface.cpp:19-82), Real = double: RA/Dec [deg] -> kR
int iph_model(b200rt_ctx *c, double g_lya, const double *marspos, int n_los, const double *ra, const double *dec,
              double *iph_kR) {
  if (!marspos || !ra || !dec || !iph_kR || n_los <= 0) return fail(c, B200RT_ERR_ARG, "b200rt_iph_model: bad argument");
  const double line_f_coeff = 2.647e-2, lyman_alpha_f = 0.41641, clight = 3e10, lyman_alpha_lambda = 121.6e-7;   // constants.hpp:22-32
  const double lyman_alpha_cross_section_total = line_f_coeff * lyman_alpha_f;
  double Fsun = g_lya / lyman_alpha_cross_section_total;
  Fsun *= (marspos[0] * marspos[0] + marspos[1] * marspos[1] + marspos[2] * marspos[2]);
  Fsun *= clight / lyman_alpha_lambda / lyman_alpha_lambda / 1e8;
  std::vector<float> u(n_los), v(n_los), w(n_los), out(n_los);
  for (int i = 0; i < n_los; i++) {
    const double thisdec = M_PI / 180 * dec[i], thisra = M_PI / 180 * ra[i];
    const double j0 = cos(thisdec) * cos(thisra), j1 = cos(thisdec) * sin(thisra), j2 = sin(thisdec);
    const double eob = M_PI / 180. * 23.44;
    u[i] = (float) j0;
    v[i] = (float) (j1 * cos(-eob) - j2 * sin(-eob));
    w[i] = (float) (j2 * cos(-eob) + j1 * sin(-eob));
  }
  if (int rc = iph_background(c, (float) Fsun, (float) marspos[0], (float) marspos[1], (float) marspos[2], n_los, u.data(),
                              v.data(), w.data(), out.data(), nullptr))
    return rc;
  for (int i = 0; i < n_los; i++) iph_kR[i] = (double) out[i] / 1000.;
  return B200RT_OK;
}

} // namespace b200rt
