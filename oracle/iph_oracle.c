/* iph_oracle.c -- TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's CPU legs).
 *
 * CPU restatement of the Quemerais interplanetary-hydrogen Lyman-alpha background model that the
 * reference calls per line of sight:
 *     quemerais_iph_model            src/quemerais_IPH_model/iph_model_interface.cpp:19-82
 *     BACKGROUND / INTENSM_PH / TOP / DEN / IPAL3M / T
 *                                    src/quemerais_IPH_model/ipbackgroundCFR_fun.f
 * The Fortran is REAL*4 throughout and so is this file (every expression is evaluated in float,
 * -ffp-contract=off).  Only what feeds the returned value xsn(2) = FLN(2) (ipbackgroundCFR_fun.f:318)
 * is evaluated: the step length comes from the idb = 1 density, the optical depth from TOP(..., idb)
 * scaled by DINF(2)/DINF(1), the source functions from SN(:,:,2), SO(:,:,2).
 *
 * PARITY UNPINNED: there is no Fortran compiler in the build container and the reference has no
 * known-answer vector for this model (python/test/obs_fit_test.cpp:60-71 only exercises it), so this
 * restatement could not be checked against the Fortran; the table-file layout was checked token for
 * token against the READ sequence (:107-164).
 *
 * Documented deviations (both also made by the CUDA kernel, csrc/iph.cu):
 *  - ACOS is evaluated by a fixed float polynomial (iph_acosf, <= 2 ulp from libm) so that the march --
 *    whose step counts are decided by float comparisons -- takes the same path on the CPU and on the
 *    GPU; its argument is clamped to [-1, 1] (the Fortran would return NaN for |y/r| > 1);
 *  - where the Fortran's bracket searches would read one element past a table with a zero weight
 *    (T == ANG(LMAX), Z == ALT(KMAX)) the index is clamped; T > ANG(LMAX) extrapolates the last interval.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NR 60
#define NK 19
#define NINF 5

typedef struct {
  int kmax, lmax, ninf;
  float alt[NR];             /* metres after scaling (ALT = ZALT) */
  float ang[NK];             /* degrees */
  float dans[NR][NK];        /* density / density at infinity */
  float sot[NR][NK];
  float so[NINF][NR][NK], sn[NINF][NR][NK];
  float dinf[NINF];          /* m^-3 after scaling */
  float alt_au[NR], dinf_cm3[NINF];   /* as parsed */
  float temp;
  /* constants of BACKGROUND (:176-221) */
  float ua, dpi, sig, dtap, sigmaf;
  float a11, a12, a13, a21, a22, a23, a31, a32, a33;
} iph_model;

static void iph_constants(iph_model *m) {
  /* ipbackgroundCFR_fun.f:190-235, IJKL = 1 (Lyman alpha) */
  const float XLA = 1.21566E-05f, PTF = 0.4162f;
  const float AM = 1.67333E-27f, BOLK = 1.38046E-23f;
  const float PY = 4 * atanf(1.f);
  m->dpi = PY / 180.f;
  const float SPI = sqrtf(PY);
  const float XNUZ = 1.f / XLA;
  const float E2 = 23.0677E-20f, EMAS = 9.1084E-28f;
  const float C = 2.99793E+10f;
  const float DLDN = XLA * XLA * 1.E+08f / C;
  const float SIGMAN = PY * E2 * PTF / (EMAS * C);
  m->sigmaf = SIGMAN * DLDN;
  const float DELNUD = XNUZ * sqrtf(2.f * BOLK * m->temp / AM) * 1.E+2f;
  float SIG = SIGMAN / (SPI * DELNUD);
  SIG = SIG * 1.E-4f;
  m->sig = SIG;
  m->dtap = 1.f / SIG;
  const float ALAMVENT = 252.3f * m->dpi, DECVENT = 8.7f * m->dpi;
  m->a11 = sinf(ALAMVENT);
  m->a12 = -cosf(ALAMVENT);
  m->a13 = 0.f;
  m->a21 = cosf(DECVENT) * cosf(ALAMVENT);
  m->a22 = cosf(DECVENT) * sinf(ALAMVENT);
  m->a23 = sinf(DECVENT);
  m->a31 = -sinf(DECVENT) * cosf(ALAMVENT);
  m->a32 = -sinf(DECVENT) * sinf(ALAMVENT);
  m->a33 = cosf(DECVENT);
}

/* arrays as parsed from the file: alt in AU, dinf in cm^-3; [k][l] and [ii][k][l] row major */
void *iph_oracle_create(int kmax, int lmax, int ninf, float temp, const float *alt_au, const float *ang,
                        const float *dans, const float *sot, const float *so, const float *sn,
                        const float *dinf_cm3) {
  if (kmax < 2 || kmax > NR - 1 || lmax < 2 || lmax > NK || ninf < 2 || ninf > NINF) return NULL;
  iph_model *m = (iph_model *) calloc(1, sizeof(iph_model));
  m->kmax = kmax; m->lmax = lmax; m->ninf = ninf; m->temp = temp;
  m->ua = 1.4959E+11f;
  for (int i = 0; i < ninf; i++) { m->dinf_cm3[i] = dinf_cm3[i]; m->dinf[i] = dinf_cm3[i] * 1.E6f; }
  for (int k = 0; k < kmax; k++) { m->alt_au[k] = alt_au[k]; m->alt[k] = alt_au[k] * m->ua; }
  for (int l = 0; l < lmax; l++) m->ang[l] = ang[l];
  for (int k = 0; k < kmax; k++)
    for (int l = 0; l < lmax; l++) {
      m->dans[k][l] = dans[k * lmax + l];
      m->sot[k][l] = sot[k * lmax + l];
      for (int i = 0; i < ninf; i++) {
        m->so[i][k][l] = so[((size_t) i * kmax + k) * lmax + l];
        m->sn[i][k][l] = sn[((size_t) i * kmax + k) * lmax + l];
      }
    }
  iph_constants(m);
  return m;
}

/* the READ sequence of BACKGROUND (:107-164): list-directed, i.e. a stream of numeric tokens */
void *iph_oracle_create_from_file(const char *fname) {
  FILE *f = fopen(fname, "r");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  char *buf = (char *) malloc(sz + 1);
  if (fread(buf, 1, sz, f) != (size_t) sz) { fclose(f); free(buf); return NULL; }
  buf[sz] = 0;
  fclose(f);
  char *p = buf;
#define NEXT() strtof(p, &p)
  int kmax = (int) NEXT(), lmax = (int) NEXT(), ninf = (int) NEXT();
  if (kmax < 2 || kmax > NR - 1 || lmax != NK || ninf < 2 || ninf > NINF) { free(buf); return NULL; }
  static float alt[NR], ang[NK], dans[NR * NK], sot[NR * NK], so[NINF * NR * NK], sn[NINF * NR * NK], dinf[NINF];
  float hdr[8], temp = 0;
  for (int i = 0; i < 8; i++) hdr[i] = NEXT();
  temp = hdr[3];
  dinf[0] = hdr[7];
  const int c0[4] = {0, 5, 10, 15}, c1[4] = {5, 10, 15, 19};
  float *first[4] = {dans, sot, so, sn};
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      for (int l = c0[b]; l < c1[b]; l++) ang[l] = NEXT();
      for (int k = 0; k < kmax; k++) {
        alt[k] = NEXT();
        for (int l = c0[b]; l < c1[b]; l++) first[a][k * lmax + l] = NEXT();
      }
    }
  for (int ii = 1; ii < ninf; ii++) {
    for (int i = 0; i < 8; i++) hdr[i] = NEXT();
    temp = hdr[3];
    dinf[ii] = hdr[7];
    float *arr[2] = {so + (size_t) ii * kmax * lmax, sn + (size_t) ii * kmax * lmax};
    for (int a = 0; a < 2; a++)
      for (int b = 0; b < 4; b++) {
        for (int l = c0[b]; l < c1[b]; l++) ang[l] = NEXT();
        for (int k = 0; k < kmax; k++) {
          (void) NEXT();   /* ZALT: overwritten by ALT*UA afterwards (:181) */
          for (int l = c0[b]; l < c1[b]; l++) arr[a][k * lmax + l] = NEXT();
        }
      }
  }
#undef NEXT
  free(buf);
  return iph_oracle_create(kmax, lmax, ninf, temp, alt, ang, dans, sot, so, sn, dinf);
}

void iph_oracle_destroy(void *h) { free(h); }

/* export the parsed tables in the layout iph_oracle_create takes (alt in AU, dinf in cm^-3) */
void iph_oracle_get_table(void *h, int *dims, float *temp, float *alt_au, float *ang, float *dans, float *sot,
                          float *so, float *sn, float *dinf_cm3) {
  iph_model *m = (iph_model *) h;
  dims[0] = m->kmax; dims[1] = m->lmax; dims[2] = m->ninf;
  *temp = m->temp;
  for (int i = 0; i < m->ninf; i++) dinf_cm3[i] = m->dinf_cm3[i];
  for (int k = 0; k < m->kmax; k++) alt_au[k] = m->alt_au[k];
  for (int l = 0; l < m->lmax; l++) ang[l] = m->ang[l];
  for (int k = 0; k < m->kmax; k++)
    for (int l = 0; l < m->lmax; l++) {
      dans[k * m->lmax + l] = m->dans[k][l];
      sot[k * m->lmax + l] = m->sot[k][l];
      for (int i = 0; i < m->ninf; i++) {
        so[((size_t) i * m->kmax + k) * m->lmax + l] = m->so[i][k][l];
        sn[((size_t) i * m->kmax + k) * m->lmax + l] = m->sn[i][k][l];
      }
    }
}

/* ---- fixed float polynomial for acos (Cephes-style asin kernel), evaluated without contraction */
static float iph_asin_core(float a) {   /* 0 <= a <= 0.5 */
  const float z = a * a;
  float p = 4.2163199048E-2f;
  p = p * z + 2.4181311049E-2f;
  p = p * z + 4.5470025998E-2f;
  p = p * z + 7.4953002686E-2f;
  p = p * z + 1.6666752422E-1f;
  return p * z * a + a;
}
static float iph_acosf(float x) {
  const float PIO2 = 1.5707963267948966f, PI_F = 3.14159265358979f;
  if (x > 1.f) x = 1.f;
  if (x < -1.f) x = -1.f;
  if (x > 0.5f) return 2.f * iph_asin_core(sqrtf(0.5f * (1.f - x)));
  if (x < -0.5f) return PI_F - 2.f * iph_asin_core(sqrtf(0.5f * (1.f + x)));
  if (x >= 0.f) return PIO2 - iph_asin_core(x);
  return PIO2 + iph_asin_core(-x);
}

/* bracket searches of DEN / IPAL3M (:517-539, 697-722): first J with T <= ANG(J) */
static void iph_bracket_ang(const iph_model *m, float T, int *ll, int *llp, float *dt) {
  for (int j = 0; j < m->lmax; j++) {
    const float d = T - m->ang[j];
    if (d < 0.f) {            /* label 22 */
      *ll = j - 1; *llp = j;
      if (j == 0) { *ll = 0; *llp = 1; }   /* T < ANG(1) cannot happen for T = acos/DPI >= 0 */
      *dt = (T - m->ang[*ll]) / (m->ang[*llp] - m->ang[*ll]);
      return;
    }
    if (d == 0.f) {           /* label 21 */
      *ll = j; *llp = (j + 1 < m->lmax) ? j + 1 : j;
      *dt = 0.f;
      return;
    }
  }
  *ll = m->lmax - 2; *llp = m->lmax - 1;
  *dt = (T - m->ang[*ll]) / (m->ang[*llp] - m->ang[*ll]);
}
static void iph_bracket_alt(const iph_model *m, float Z, int *kk, int *kkp, float *du) {
  for (int k = 0; k < m->kmax; k++) {
    const float d = Z - m->alt[k];
    if (d < 0.f) {            /* label 26 (k >= 1 because Z >= ALT(1)) */
      *kk = k - 1; *kkp = k;
      *du = (Z - m->alt[*kk]) / (m->alt[*kkp] - m->alt[*kk]);
      return;
    }
    if (d == 0.f) {           /* label 25 */
      *kk = k; *kkp = (k + 1 < m->kmax) ? k + 1 : k;
      *du = 0.f;
      return;
    }
  }
  *kk = m->kmax - 1; *kkp = m->kmax - 1; *du = 0.f;   /* unreachable: Z is clamped to ALT(KMAX) */
}

/* DEN (:492-547): relative density at (Z, T); *ko = 1-based index of the lower radial node */
static float iph_den(const iph_model *m, float Z, float T, int *ko) {
  *ko = 1;
  if (Z < m->alt[0]) return 0.f;
  if (Z > m->alt[m->kmax - 1]) Z = m->alt[m->kmax - 1];
  int ll, llp, kk, kkp;
  float dt, du;
  iph_bracket_ang(m, T, &ll, &llp, &dt);
  iph_bracket_alt(m, Z, &kk, &kkp, &du);
  const float fl = m->dans[kk][ll] + du * (m->dans[kkp][ll] - m->dans[kk][ll]);
  const float flp = m->dans[kk][llp] + du * (m->dans[kkp][llp] - m->dans[kk][llp]);
  *ko = kk + 1;
  return fl + dt * (flp - fl);
}

/* IPAL3M (:659-741) for one density index */
static void iph_ipal3m(const iph_model *m, float R, float T, int imd, float *foo, float *fn) {
  *fn = 0.f;     /* F and CT are zeroed before the early return, FOO is NOT (:685-690): the caller's value survives */
  if (R < m->alt[0] || R >= m->alt[m->kmax - 1]) return;
  int ll, llp, kk, kkp;
  float dt, du;
  iph_bracket_ang(m, T, &ll, &llp, &dt);
  iph_bracket_alt(m, R, &kk, &kkp, &du);
  float fl = m->sn[imd][kk][ll] + du * (m->sn[imd][kkp][ll] - m->sn[imd][kk][ll]);
  float flp = m->sn[imd][kk][llp] + du * (m->sn[imd][kkp][llp] - m->sn[imd][kk][llp]);
  *fn = fl + dt * (flp - fl);
  fl = m->so[imd][kk][ll] + du * (m->so[imd][kkp][ll] - m->so[imd][kk][ll]);
  flp = m->so[imd][kk][llp] + du * (m->so[imd][kkp][llp] - m->so[imd][kk][llp]);
  *foo = fl + dt * (flp - fl);
}

/* T (:363-397): Holstein transmission */
static float iph_T(float TO) {
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

/* TOP (:399-490): optical depth between two points, density index imd */
static float iph_top(const iph_model *m, float XF, float YF, float ZF, float XH, float YH, float ZH, int imd) {
  const float UA = m->ua;
  float XA = XF / UA, XB = XH / UA, YA = YF / UA, YB = YH / UA, ZA = ZF / UA, ZB = ZH / UA;
  const float altp = m->alt[0] / UA;
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
  float TA = iph_acosf(YA / RA) / m->dpi;
  int KP;
  float DN1 = iph_den(m, RA * UA, TA, &KP);
  DN1 = m->dinf[imd] * DN1;
  if (KP == m->kmax) KP = KP - 1;
  float DMA = (m->alt[KP] - m->alt[KP - 1]) / 3.f / UA;
  float DSAB = fminf(DMA, DSA0);
  float SAB = 0.f, DT = 0.f;
  do {
    XA = XA + DSAB * XAB;
    YA = YA + DSAB * YAB;
    ZA = ZA + DSAB * ZAB;
    SAB = SAB + DSAB;
    RA = sqrtf(XA * XA + YA * YA + ZA * ZA);
    TA = iph_acosf(YA / RA) / m->dpi;
    float DN = iph_den(m, RA * UA, TA, &KP);
    DN = m->dinf[imd] * DN;
    DT = DT + (DN + DN1) * .5f * DSAB * m->sig * UA;
    DN1 = DN;
    if (KP == m->kmax) KP = KP - 1;
    DMA = (m->alt[KP] - m->alt[KP - 1]) / 3.f / UA;
    DSAB = fminf(DMA, DSA0);
  } while (SAB <= NORME);
  return DT;
}

/* INTENSM_PH (:550-657), value FLN(2) only.  n_steps (may be NULL) returns the outer step count */
static float iph_intens(const iph_model *m, float GRAL, float X, float Y, float Z, float U, float V, float W,
                        int *n_steps) {
  const int idb = 0, iout = 1;
  const float UA = m->ua;
  float S = 0.f, TT = 0.f, FLN = 0.f;
  if (n_steps) *n_steps = 0;
  const float RR = sqrtf(X * X + Y * Y + Z * Z);
  if (RR > m->alt[m->kmax - 1]) return 0.f;
  float YP = Y, R = RR, XAV = X, YAV = Y, ZAV = Z;
  float FOO = 0.f;   /* FOO(5) is zeroed once per line of sight (:577-585) and then only written by IPAL3M */
  for (;;) {
    const float TETA = iph_acosf(YP / R) / m->dpi;
    int KO;
    const float DNA = iph_den(m, R, TETA, &KO);
    float DN1 = m->dinf[idb] * DNA;
    if (DN1 == 0.f) DN1 = 1.f;
    float DP = m->dtap * 0.05f / DN1;
    float DUA;
    if (KO < m->kmax) DUA = (m->alt[KO] - m->alt[KO - 1]) / 2.f;
    else DUA = (m->alt[m->kmax - 1] - m->alt[m->kmax - 2]) / 2.f;
    DP = fminf(DP, DUA);
    DP = fmaxf(DP, UA / 10.f);
    S = S + DP;
    const float XP = X + S * U;
    YP = Y + S * V;
    const float ZP = Z + S * W;
    R = sqrtf(XP * XP + YP * YP + ZP * ZP);
    if (R > m->alt[m->kmax - 1]) break;
    const float TETA2 = iph_acosf(YP / R) / m->dpi;
    float FN;
    iph_ipal3m(m, R, TETA2, iout, &FOO, &FN);
    const float DTT = iph_top(m, XAV, YAV, ZAV, XP, YP, ZP, idb);
    const float cosff = (U * XP + V * YP + W * ZP) / R;
    const float corec = 0.25f * cosff * cosff + (11.f / 12.f);
    TT = TT + DTT * m->dinf[iout] / m->dinf[idb];
    const float FFNN = FN + FOO * (corec - 1.f);
    const float TTTII = iph_T(TT);
    const float DFLNC = FFNN * GRAL * TTTII * DP;
    FLN = FLN + DFLNC;
    XAV = XP; YAV = YP; ZAV = ZP;
    if (n_steps) (*n_steps)++;
  }
  return FLN;
}

/* BACKGROUND (:1-324): fs = solar line-centre flux at 1 AU [ph cm^-2 s^-1 A^-1], observer position [AU,
 * ecliptic], look directions (ecliptic unit vectors); fln = xsn(2) [R] */
void iph_oracle_background(void *h, float fs, float xpos, float ypos, float zpos, int n_los, const float *u1,
                           const float *v1, const float *w1, float *fln, int *n_steps) {
  const iph_model *m = (const iph_model *) h;
  const float GZERO = fs * m->sigmaf;
  const float GRAL = GZERO * 1.E-10f;
  const float x2 = (m->a11 * xpos + m->a12 * ypos) * m->ua;
  const float y2 = (m->a21 * xpos + m->a22 * ypos + m->a23 * zpos) * m->ua;
  const float z2 = (m->a31 * xpos + m->a32 * ypos + m->a33 * zpos) * m->ua;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n_los; i++) {
    const float u2 = m->a11 * u1[i] + m->a12 * v1[i];
    const float v2 = m->a21 * u1[i] + m->a22 * v1[i] + m->a23 * w1[i];
    const float w2 = m->a31 * u1[i] + m->a32 * v1[i] + m->a33 * w1[i];
    fln[i] = iph_intens(m, GRAL, x2, y2, z2, u2, v2, w2, n_steps ? n_steps + i : NULL);
  }
}

/* quemerais_iph_model (iph_model_interface.cpp:19-82), Real = double: RA/Dec [deg] -> kR */
void iph_oracle_model(void *h, double g_lya, const double *marspos, int n_los, const double *ra, const double *dec,
                      double *iph_kR) {
  const double line_f_coeff = 2.647e-2, lyman_alpha_f = 0.41641, clight = 3e10, lyman_alpha_lambda = 121.6e-7;
  const double lyman_alpha_cross_section_total = line_f_coeff * lyman_alpha_f;
  double Fsun = g_lya / lyman_alpha_cross_section_total;
  Fsun *= (marspos[0] * marspos[0] + marspos[1] * marspos[1] + marspos[2] * marspos[2]);
  Fsun *= clight / lyman_alpha_lambda / lyman_alpha_lambda / 1e8;
  float *u = (float *) malloc(sizeof(float) * 4 * (n_los > 0 ? n_los : 1));
  float *v = u + n_los, *w = v + n_los, *out = w + n_los;
  for (int i = 0; i < n_los; i++) {
    const double thisdec = M_PI / 180 * dec[i], thisra = M_PI / 180 * ra[i];
    const double j0 = cos(thisdec) * cos(thisra), j1 = cos(thisdec) * sin(thisra), j2 = sin(thisdec);
    const double eob = M_PI / 180. * 23.44;
    u[i] = (float) j0;
    v[i] = (float) (j1 * cos(-eob) - j2 * sin(-eob));
    w[i] = (float) (j2 * cos(-eob) + j1 * sin(-eob));
  }
  iph_oracle_background(h, (float) Fsun, (float) marspos[0], (float) marspos[1], (float) marspos[2], n_los, u, v, w,
                        out, NULL);
  for (int i = 0; i < n_los; i++) iph_kR[i] = (double) out[i] / 1000.;   /* iphb_[i]/1000. with Real = double */
  free(u);
}
