/* multiplet_oracle.inc.c -- CPU restatement of the reference's multiplet CFR emissions.
 * TEST INFRASTRUCTURE ONLY; #included at the end of rt_oracle.c (shares its grid, traversal and
 * line-of-sight geometry).  Paths relative to /root/reference/src.
 *
 *   multiplet_CFR_emission::update_tracker_start / update_tracker_influence / update_tracker_brightness
 *                                       emission/multiplet_CFR_emission.hpp:69-300, 382-404, 302-316
 *   O_1026_tracker / O_1026_emission    emission/O_1026_tracker.hpp:11-291, emission/O_1026.hpp:84-217
 *   H_lyman_multiplet(_tracker)         emission/H_multiplet_tracker.hpp:11-248, emission/H_lyman_multiplet.hpp:118-217
 *   H_lyman_singlet(_tracker)           emission/H_multiplet_tracker_test.hpp, emission/H_lyman_multiplet_test.hpp
 *   emission_voxels (element order, accumulate_influence, solve, brightness interp)
 *                                       emission/emission_voxels.hpp:58-70, 137-155, 170-176, 199-233
 *
 * PINNING: tests/test_multiplet.py checks this against the reference's own source compiled in place
 * (oracle/_ref/libref_mult_f{64,32}.so: constants and line shapes bit for bit, K / S0 / brightness bit for
 * bit) and tests/golden/mult_*.npz holds fixtures made from that build.
 *
 * Type rule reproduced for Real = float: grid/coordinate_generation.hpp:19-20 puts `using std::exp; using std::log;`
 * at global scope and is included before the emission headers (observation_fit.hpp:15-21), so the trackers'
 * unqualified exp() is the float overload, while their unqualified sqrt() is the double libm function
 * (rounded on assignment).
 */
#define M_MAXL 6
#define M_MAXM 3
#define M_MAXLOW 3
#define M_MAXUP 4
#define M_MAXLAM 41


typedef struct {
  int kind, n_lines, n_mult, n_lower, n_upper, n_lambda;
  int mult[M_MAXL], lower[M_MAXL], upper[M_MAXL];
  REAL sigma[M_MAXL], A[M_MAXL], xsec[M_MAXL], decay[M_MAXL];   /* decay is indexed by UPPER state */
  REAL offset[M_MAXL];    /* line_wavelength_offset / doppler_width_wavelength_reference (dimensionless) */
  REAL norm[M_MAXL];      /* line shape normalisation at T_ref [1/Hz] */
  REAL weight[M_MAXL];    /* delta_lambda * doppler_width_frequency_reference [Hz] */
  REAL T_ref, lambda_max, delta_lambda;
  REAL solar[M_MAXL];     /* pumping flux of each line [ph/cm2/s/Hz] */
  int pumped[M_MAXL];     /* 1 = compute_single_scattering assigns singlescat from this line */
  REAL lower_energy[M_MAXLOW], lower_g[M_MAXLOW];   /* O I only (Boltzmann populations) */
} omult;

typedef struct {
  omult d;
  REAL *n[M_MAXLOW], *n_pt[M_MAXLOW], *T, *T_pt, *nabs, *nabs_pt;
  REAL *K, *S0, *tsp, *tab, *S;
} omstate;

/* Real.hpp:63-82 */
static REAL om_constexpr_sqrt(REAL x) {
  REAL curr=x, prev=0;
  while (curr != prev) { REAL next = 0.5*(curr + x/curr); prev=curr; curr=next; }
  return curr;
}

/* constants.hpp */
static const REAL om_kB = RL(1.38e-16), om_erg_per_eV = RL(1.60218e-12), om_clight = RL(3e10), om_mH = RL(1.673e-24);
static const REAL om_line_f_coeff = RL(2.647e-2);

/* doppler-width block shared by the three trackers (O_1026_tracker.hpp:150-160, H_multiplet_tracker.hpp:100-124):
   given the reference wavelength [nm] and velocity, the wavelength width, frequency width and normalisation */
static void om_doppler(REAL ref_lambda, REAL ref_velocity, REAL *waveref, REAL *freqref, REAL *normalization) {
  *waveref = (ref_lambda*ref_velocity/om_clight);
  *freqref = (REAL) (1.0/(ref_lambda*1e-7)*ref_velocity);
  *normalization = (o_one_over_sqrt_pi / *freqref);
}

static void om_desc(int kind, omult *d) {
  memset(d, 0, sizeof(*d));
  d->kind=kind;
  d->T_ref=200; d->lambda_max=RL(4.0);
  if (kind==0) {   /* O_1026_constants_detail + O_1026_tracker */
    static const int mi[6]={0,1,1,2,2,2}, li[6]={0,1,1,2,2,2}, ui[6]={0,0,1,0,1,2}, lowJ[6]={0,1,1,2,2,2};
    static const double wl[6]={102.81571,102.74313,102.74305,102.57633,102.57626,102.57616};
    static const double off[6]={0.0,4e-5,-4e-5,8e-5,1e-5,-9e-5};
    static const double A[6]={4.22e7,3.17e7,5.71e7,2.11e6,1.91e7,7.66e7};
    static const double f[6]={2.01e-2,5.02e-3,1.51e-2,2.00e-4,3.01e-3,1.69e-2};
    d->n_lines=6; d->n_mult=3; d->n_lower=3; d->n_upper=3; d->n_lambda=21;
    REAL vel = om_constexpr_sqrt(2*om_kB*d->T_ref/(16*om_mH));
    REAL waveref, freqref, normalization;
    om_doppler((REAL) wl[5], vel, &waveref, &freqref, &normalization);
    d->delta_lambda = 2*d->lambda_max/(d->n_lambda-1);
    for (int l=0;l<6;l++) {
      d->mult[l]=mi[l]; d->lower[l]=li[l]; d->upper[l]=ui[l];
      d->A[l]=(REAL) A[l]; d->sigma[l]=om_line_f_coeff*(REAL) f[l]; d->xsec[l]=(REAL) 3.53e-17;
      d->offset[l]=(REAL) off[l]/waveref; d->norm[l]=normalization; d->weight[l]=d->delta_lambda*freqref;
      d->pumped[l]=(lowJ[l]==2);
    }
    d->decay[0]=(REAL) (2.11e6 + 3.17e7 + 4.22e7 + 1.29e7 + 8.6e5 + 1.72e7);
    d->decay[1]=(REAL) (1.91e7 + 5.71e7 + 2.32e7 + 7.74e6);
    d->decay[2]=(REAL) (7.66e7 + 3.09e7);
    d->lower_energy[0]=RL(0.0281416)*om_erg_per_eV; d->lower_energy[1]=RL(0.0196224)*om_erg_per_eV;
    d->lower_energy[2]=RL(0.0)*om_erg_per_eV;
    d->lower_g[0]=1; d->lower_g[1]=3; d->lower_g[2]=5;
  } else {
    const int single = (kind==2);
    static const int mi4[4]={0,0,1,1};
    static const double wl4[4]={121.5668237310,121.5673644608,102.572182505,102.572296565};
    static const double off4[4]={-2.70365e-4,2.70365e-4,-5.703e-5,5.703e-5};
    static const double A4[4]={6.2648e8,6.2649e8,1.6725e8,1.6725e8};
    static const double f4[4]={0.2776,0.13881,5.2761e-2,2.6381e-2};
    static const double x4[4]={6.3e-20,6.3e-20,3.53e-17,3.52e-17};
    const double lya_nm = 121.6e-7*1e7, lyb_nm = 102.6e-7*1e7;   /* lyman_alpha_lambda*1e7 */
    d->n_lines = single ? 2 : 4; d->n_mult=2; d->n_lower=1; d->n_upper = single ? 2 : 4; d->n_lambda=41;
    REAL vel = om_constexpr_sqrt(2*om_kB*d->T_ref/om_mH);
    REAL wr[2], fr[2], nm[2];
    if (single) {
      /* line_wavelength = {lyman_alpha_lambda*1e7, lyman_beta_lambda*1e7}: Real * double, rounded to Real */
      REAL la = (REAL) ((REAL) 121.6e-7*1e7), lb = (REAL) ((REAL) 102.6e-7*1e7);
      (void) lya_nm; (void) lyb_nm;
      om_doppler(la, vel, &wr[0], &fr[0], &nm[0]);
      om_doppler(lb, vel, &wr[1], &fr[1], &nm[1]);
    } else {
      om_doppler((REAL) wl4[0], vel, &wr[0], &fr[0], &nm[0]);
      om_doppler((REAL) wl4[2], vel, &wr[1], &fr[1], &nm[1]);
    }
    d->delta_lambda = 2*d->lambda_max/(d->n_lambda-1);
    for (int l=0;l<d->n_lines;l++) {
      const int g = single ? l : (l<2 ? 0 : 1);       /* Lyman alpha or beta */
      d->mult[l] = single ? l : mi4[l]; d->lower[l]=0; d->upper[l]=l;
      if (single) {
        d->A[l]=(REAL) (l==0 ? 6.2648e8 : 1.6725e8);
        d->sigma[l]=om_line_f_coeff*(REAL) (l==0 ? 0.2776 + 0.13881 : 5.2761e-2 + 2.6381e-2);
        d->xsec[l]=(REAL) (l==0 ? 6.3e-20 : 3.52e-17);
        d->offset[l]=(REAL) 0.0/wr[g];
        d->decay[l]=(REAL) (l==0 ? 6.2648e8 : 1.6725e8 + 2.2449e7);
      } else {
        d->A[l]=(REAL) A4[l]; d->sigma[l]=om_line_f_coeff*(REAL) f4[l]; d->xsec[l]=(REAL) x4[l];
        d->offset[l]=(REAL) off4[l]/wr[g];
        d->decay[l]=(REAL) (l==0 ? 6.2648e8 : l==1 ? 6.2649e8 : 1.6725e8 + 2.2449e7);
      }
      d->norm[l]=nm[g]; d->weight[l]=d->delta_lambda*fr[g];
      d->pumped[l]=1;
    }
  }
}

/* tracker statics: lambda(), line_shape_function, line_shape_normalization (O_1026_tracker.hpp:199-237) */
static REAL om_lambda(const omult *d, int i) { return (-d->lambda_max + i*d->delta_lambda); }
static REAL om_shape(const omult *d, int line, int i, REAL T) {
  REAL lambda2 = (om_lambda(d,i) - d->offset[line]);
  lambda2 *= lambda2;
  lambda2 = lambda2*d->T_ref/T;
  lambda2 = STD_EXP(-lambda2);
  return lambda2;
}
static REAL om_normT(const omult *d, int line, REAL T) {
  return (REAL) (d->norm[line]*sqrt(d->T_ref/T));
}
static REAL om_shape_norm(const omult *d, int line, int i, REAL T) { return om_normT(d,line,T)*om_shape(d,line,i,T); }

typedef struct {
  REAL col[M_MAXLOW];
  REAL tau_sp[M_MAXL], tau_ab[M_MAXL];
  REAL T_final[M_MAXL], T_int[M_MAXL], G[M_MAXUP][M_MAXUP], brightness[M_MAXL];
  REAL T0, n0[M_MAXLOW];
  REAL P[M_MAXM][M_MAXLAM];
} omtracker;

static void om_reset(const omult *d, omtracker *t, REAL T0, const REAL *n0) {   /* tracker.reset */
  t->T0=T0;
  for (int l=0;l<d->n_lines;l++) { t->tau_sp[l]=0.0; t->tau_ab[l]=0.0; t->brightness[l]=0.0; }
  for (int m=0;m<d->n_mult;m++) for (int i=0;i<d->n_lambda;i++) t->P[m][i]=1.0;
  for (int l=0;l<d->n_lower;l++) { t->n0[l]=n0 ? n0[l] : 0; t->col[l]=0.0; }
}

/* multiplet_CFR_emission::update_tracker_start<influence>, multiplet_CFR_emission.hpp:69-300 */
static void om_update(const omult *d, int influence, REAL T, const REAL *dens, REAL nabs, REAL pathlength, omtracker *t) {
  REAL lineshape[M_MAXL][M_MAXLAM], tau_lambda[M_MAXM][M_MAXLAM], tp_voxel[M_MAXM][M_MAXLAM], tp_final[M_MAXM][M_MAXLAM];
  REAL coefa[M_MAXL][M_MAXLAM], ls0[M_MAXL][M_MAXLAM];
  const int NL=d->n_lines, NM=d->n_mult, NLAM=d->n_lambda;
  for (int l=0;l<d->n_lower;l++) t->col[l] += dens[l]*pathlength;
  for (int m=0;m<NM;m++) for (int i=0;i<NLAM;i++) tau_lambda[m][i]=0.0;
  for (int line=0;line<NL;line++) {
    const int lo=d->lower[line], m=d->mult[line];
    for (int i=0;i<NLAM;i++) {
      lineshape[line][i] = om_shape_norm(d,line,i,T);
      tau_lambda[m][i] += ((dens[lo]*d->sigma[line]*lineshape[line][i] + nabs*d->xsec[line])*pathlength);
    }
    REAL tsv = ((dens[lo]*d->sigma[line]*om_normT(d,line,T))*pathlength);
    t->tau_sp[line] += tsv;
    REAL tav = (nabs*d->xsec[line]*pathlength);
    t->tau_ab[line] += tav;
  }
  for (int m=0;m<NM;m++)
    for (int i=0;i<NLAM;i++) {
      tp_voxel[m][i] = STD_EXP(-tau_lambda[m][i]);
      tp_final[m][i] = (t->P[m][i]*tp_voxel[m][i]);
    }
  for (int line=0;line<NL;line++) { t->T_int[line]=0; t->T_final[line]=0; }
  for (int a=0;a<d->n_upper;a++) for (int b=0;b<d->n_upper;b++) t->G[a][b]=0;
  for (int line=0;line<NL;line++) {
    const int m=d->mult[line];
    for (int i=0;i<NLAM;i++) {
      REAL holcoef = d->weight[line];
      if (tau_lambda[m][i] < 1e-3) coefa[line][i] = (RL(1.0) - (RL(0.5)*tau_lambda[m][i]));
      else coefa[line][i] = ((RL(1.0) - tp_voxel[m][i])/(tau_lambda[m][i]));
      coefa[line][i] *= (holcoef*lineshape[line][i]*t->P[m][i]*pathlength);
      t->T_int[line] += coefa[line][i];
      if (influence) {
	ls0[line][i] = om_shape_norm(d,line,i,t->T0);
	t->T_final[line] += (holcoef*ls0[line][i]*tp_final[m][i]);
      }
    }
  }
  if (influence) {
    for (int lo=0;lo<NL;lo++)
      for (int lc=0;lc<NL;lc++)
	if (d->mult[lo]==d->mult[lc])
	  for (int i=0;i<NLAM;i++)
	    t->G[d->upper[lo]][d->upper[lc]] += (d->sigma[lo]*t->n0[d->lower[lo]]*d->A[lc]/d->decay[d->upper[lo]]
						   *ls0[lo][i]*coefa[lc][i]);
  }
  for (int m=0;m<NM;m++) for (int i=0;i<NLAM;i++) t->P[m][i]=tp_final[m][i];
  for (int line=0;line<NL;line++) if (t->T_int[line] > pathlength) t->T_int[line]=pathlength;
}

static omstate *om_get(omodel *m);   /* storage hook, defined below */

/* kind 0: O_1026_emission::define (O_1026.hpp:134-217), 1/2: H_lyman_multiplet::define (H_lyman_multiplet.hpp:160-217)
   with the default options (atmospheric temperature, CO2 absorption on).  vox_in = [6][n_vox]:
   species avg, pt; temperature avg, pt; absorber avg, pt.  solar[2]: O I: Lyman beta flux; H: Lyman alpha, beta */
void oracle_define_multiplet(void *h, int kind, const double *solar, const double *vox_in) {
  omodel *m=(omodel*) h;
  omstate *s=om_get(m);
  const int N=m->n_vox;
  om_desc(kind, &s->d);
  omult *d=&s->d;
  for (int l=0;l<d->n_lines;l++) {
    if (kind==0) d->solar[l]=(REAL) solar[0];
    else if (kind==1) d->solar[l]=(REAL) (l<2 ? solar[0] : solar[1]);
    else d->solar[l]=(REAL) (l==0 ? solar[0] : solar[1]);
  }
  const int NE=N*d->n_upper;
  for (int l=0;l<M_MAXLOW;l++) { free(s->n[l]); free(s->n_pt[l]); s->n[l]=ralloc(N); s->n_pt[l]=ralloc(N); }
  free(s->T); free(s->T_pt); free(s->nabs); free(s->nabs_pt);
  s->T=ralloc(N); s->T_pt=ralloc(N); s->nabs=ralloc(N); s->nabs_pt=ralloc(N);
  free(s->K); free(s->S0); free(s->S); free(s->tsp); free(s->tab);
  s->K=ralloc((size_t) NE*NE); s->S0=ralloc(NE); s->S=ralloc(NE);
  s->tsp=ralloc((size_t) N*d->n_lines); s->tab=ralloc((size_t) N*d->n_lines);
  for (int i=0;i<N;i++) {
    REAL bulk=(REAL) vox_in[0*N+i], bulk_pt=(REAL) vox_in[1*N+i];
    s->T[i]=(REAL) vox_in[2*N+i]; s->T_pt[i]=(REAL) vox_in[3*N+i];
    s->nabs[i]=(REAL) vox_in[4*N+i]; s->nabs_pt[i]=(REAL) vox_in[5*N+i];
    if (kind==0) {
      REAL fr[M_MAXLOW], frp[M_MAXLOW], tot=0, totp=0;
      for (int l=0;l<3;l++) {
	fr[l] = (d->lower_g[l]*STD_EXP(-d->lower_energy[l]/om_kB/s->T[i]));
	tot += fr[l];
	frp[l] = (d->lower_g[l]*STD_EXP(-d->lower_energy[l]/om_kB/s->T_pt[i]));
	totp += frp[l];
      }
      for (int l=0;l<3;l++) {
	fr[l] /= tot; s->n[l][i] = fr[l]*bulk;
	frp[l] /= totp; s->n_pt[l][i] = frp[l]*bulk_pt;
      }
    } else { s->n[0][i]=bulk; s->n_pt[0][i]=bulk_pt; }
  }
}

void oracle_multiplet_dims(void *h, int *o) {
  omodel *m=(omodel*) h; omult *d=&om_get(m)->d;
  o[0]=m->n_vox; o[1]=m->n_rays; o[2]=d->n_lines; o[3]=d->n_mult; o[4]=d->n_lower; o[5]=d->n_upper; o[6]=d->n_lambda;
}
/* same layout as refm_constants (oracle/ref_harness_multiplet.cpp) */
void oracle_multiplet_constants(void *h, int *iout, double *dout) {
  omodel *m=(omodel*) h; omult *d=&om_get(m)->d;
  const int NL=d->n_lines;
  for (int l=0;l<NL;l++) {
    iout[0*NL+l]=d->mult[l]; iout[1*NL+l]=d->lower[l]; iout[2*NL+l]=d->upper[l];
    dout[0*NL+l]=d->sigma[l]; dout[1*NL+l]=d->A[l]; dout[2*NL+l]=d->xsec[l];
    dout[3*NL+l]=(l<d->n_upper) ? (double) d->decay[l] : 0.0;
    dout[4*NL+l]=d->offset[l]; dout[5*NL+l]=om_normT(d,l,d->T_ref); dout[6*NL+l]=d->weight[l];
  }
}
double oracle_multiplet_lineshape(void *h, int line, int i_lambda, double T) {
  omodel *m=(omodel*) h; return om_shape_norm(&om_get(m)->d, line, i_lambda, (REAL) T);
}
void oracle_multiplet_get_arrays(void *h, double *out) {
  omodel *m=(omodel*) h; omstate *s=om_get(m); const int N=m->n_vox, NLOW=s->d.n_lower;
  for (int l=0;l<NLOW;l++) for (int i=0;i<N;i++) { out[(size_t)(2*l)*N+i]=s->n[l][i]; out[(size_t)(2*l+1)*N+i]=s->n_pt[l][i]; }
  double *o=out+(size_t) 2*NLOW*N;
  for (int i=0;i<N;i++) { o[i]=s->T[i]; o[N+i]=s->T_pt[i]; o[2*N+i]=s->nabs[i]; o[3*N+i]=s->nabs_pt[i]; }
}

/* RT_grid::generate_S loop body (RT_grid.hpp:166-201) for a multiplet emission: rows of source voxels v0..v1-1;
   update_tracker_influence multiplet_CFR_emission.hpp:382-404; accumulate_influence emission_voxels.hpp:137-155;
   compute_single_scattering O_1026.hpp:84-131 / H_lyman_multiplet.hpp:118-157 */
long oracle_multiplet_build_rows(void *h, int v0, int v1, int stride) {
  omodel *m=(omodel*) h; omstate *s=om_get(m); const omult *d=&s->d;
  const int N=m->n_vox, NUP=d->n_upper, NL=d->n_lines, NE=N*NUP;
  long steps=0;
#pragma omp parallel reduction(+:steps)
  {
    obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
    REAL *infl=ralloc((size_t) NUP*NE);               /* tracker.influence[iu](voxel, ju) */
    int *touched=(int*) malloc(sizeof(int)*m->cap);
    omtracker *t=(omtracker*) malloc(sizeof(omtracker));
#pragma omp for schedule(dynamic,1)
    for (int iv=v0;iv<v1;iv+=stride) {
      REAL n0[M_MAXLOW];
      for (int l=0;l<d->n_lower;l++) n0[l]=s->n[l][iv];
      for (int iu=0;iu<NUP;iu++) memset(s->K+(size_t)(iv*NUP+iu)*NE, 0, sizeof(REAL)*NE);
      for (int ir=0;ir<m->n_rays;ir++) {
	ovec v; o_ptray(m,iv,ir,&v);
	int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
	if (n==0) continue;
	om_reset(d,t,s->T[iv],n0);
	int nt=0;
	for (int k=1;k<n;k++) {
	  int vox=b[begin+k-1].entering;
	  REAL path=b[begin+k].distance-b[begin+k-1].distance;
	  REAL dens[M_MAXLOW];
	  for (int l=0;l<d->n_lower;l++) dens[l]=s->n[l][vox];
	  om_update(d,1,s->T[vox],dens,s->nabs[vox],path,t);
	  touched[nt++]=vox;
	  for (int iu=0;iu<NUP;iu++)
	    for (int ju=0;ju<NUP;ju++) {
	      REAL coef=m->ray_domega[ir];
	      coef*=t->G[iu][ju];
	      infl[(size_t)iu*NE+vox*NUP+ju]+=coef;
	    }
	  steps++;
	}
	for (int iu=0;iu<NUP;iu++) {
	  REAL *row=s->K+(size_t)(iv*NUP+iu)*NE;
	  for (int q=0;q<nt;q++) {
	    int vox=touched[q];
	    for (int ju=0;ju<NUP;ju++) {
	      REAL c=infl[(size_t)iu*NE+vox*NUP+ju];
	      if (c!=0) { row[vox*NUP+ju]+=c; infl[(size_t)iu*NE+vox*NUP+ju]=0; }
	    }
	  }
	}
      }
      /* single scattering */
      int visible = !(m->vz[iv]<0 && m->vx[iv]*m->vx[iv]+m->vy[iv]*m->vy[iv] < m->rmin*m->rmin);
      om_reset(d,t,s->T[iv],n0);
      for (int line=0;line<NL;line++) { t->T_final[line]=0; }
      if (visible) {
	ovec v; o_ptxyz(m->vx[iv],m->vy[iv],m->vz[iv], iv, m->vr[iv], m->vt[iv], 0., 0., 1., &v);
	int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
	for (int k=1;k<n;k++) {
	  int vox=b[begin+k-1].entering;
	  REAL path=b[begin+k].distance-b[begin+k-1].distance;
	  REAL dens[M_MAXLOW];
	  for (int l=0;l<d->n_lower;l++) dens[l]=s->n[l][vox];
	  om_update(d,1,s->T[vox],dens,s->nabs[vox],path,t);
	}
      }
      for (int line=0;line<NL;line++) {
	if (!visible) { t->tau_sp[line]=-1.0; t->tau_ab[line]=-1.0; t->T_final[line]=0.0; }
	s->tsp[(size_t)iv*NL+line]=t->tau_sp[line];
	s->tab[(size_t)iv*NL+line]=t->tau_ab[line];
	if (d->pumped[line]) {
	  REAL exc=(d->solar[line]*s->n[d->lower[line]][iv]*d->sigma[line]/d->decay[d->upper[line]]);
	  s->S0[iv*NUP+d->upper[line]]=exc*t->T_final[line];
	}
      }
    }
    free(b); free(infl); free(touched); free(t);
  }
  return steps;
}

/* emission_voxels::solve (emission_voxels.hpp:170-176; pre_solve is a no-op for multiplets): (I - K) S = S0.
   The reference uses Eigen PartialPivLU (third party); textbook LU with partial pivoting here, in double. */
double oracle_multiplet_solve(void *h) {
  omodel *m=(omodel*) h; omstate *s=om_get(m);
  const int N=m->n_vox*s->d.n_upper;
  double *A=(double*) malloc(sizeof(double)*(size_t)N*N);
  double *x=(double*) malloc(sizeof(double)*N);
  for (int i=0;i<N;i++) { for (int j=0;j<N;j++) A[(size_t)i*N+j]=(i==j ? 1.0 : 0.0)-(double) s->K[(size_t)i*N+j]; x[i]=s->S0[i]; }
  for (int k=0;k<N;k++) {
    int p=k; double best=fabs(A[(size_t)k*N+k]);
    for (int i=k+1;i<N;i++) { double vv=fabs(A[(size_t)i*N+k]); if (vv>best) {best=vv;p=i;} }
    if (p!=k) { for (int j=0;j<N;j++) { double tt=A[(size_t)k*N+j]; A[(size_t)k*N+j]=A[(size_t)p*N+j]; A[(size_t)p*N+j]=tt; } double tt=x[k]; x[k]=x[p]; x[p]=tt; }
    double inv=1.0/A[(size_t)k*N+k];
#pragma omp parallel for schedule(static)
    for (int i=k+1;i<N;i++) { double f=A[(size_t)i*N+k]*inv; if (f!=0) { double *ri=A+(size_t)i*N; const double *rk=A+(size_t)k*N; for (int j=k+1;j<N;j++) ri[j]-=f*rk[j]; x[i]-=f*x[k]; } }
  }
  for (int k=N-1;k>=0;k--) { double sum=x[k]; for (int j=k+1;j<N;j++) sum-=A[(size_t)k*N+j]*x[j]; x[k]=sum/A[(size_t)k*N+k]; }
  for (int i=0;i<N;i++) s->S[i]=(REAL) x[i];
  double rmax=0, smax=0;
  for (int i=0;i<N;i++) {
    double acc=0;
    for (int j=0;j<N;j++) acc += ((i==j ? 1.0 : 0.0)-(double) s->K[(size_t)i*N+j])*(double) s->S[j];
    double rr=fabs(acc-(double) s->S0[i]); if (rr>rmax) rmax=rr;
    if (fabs((double) s->S0[i])>smax) smax=fabs((double) s->S0[i]);
  }
  free(A); free(x);
  return rmax/(smax>0 ? smax : 1.0);
}

void oracle_multiplet_get_K(void *h, double *out) {
  omodel *m=(omodel*) h; omstate *s=om_get(m); size_t N=(size_t) m->n_vox*s->d.n_upper;
  for (size_t i=0;i<N*N;i++) out[i]=s->K[i];
}
void oracle_multiplet_get_vectors(void *h, double *S0, double *tsp, double *tab, double *S) {
  omodel *m=(omodel*) h; omstate *s=om_get(m);
  for (int i=0;i<m->n_vox*s->d.n_upper;i++) { S0[i]=s->S0[i]; S[i]=s->S[i]; }
  for (int i=0;i<m->n_vox*s->d.n_lines;i++) { tsp[i]=s->tsp[i]; tab[i]=s->tab[i]; }
}
void oracle_multiplet_set_sourcefn(void *h, const double *S) {
  omodel *m=(omodel*) h; omstate *s=om_get(m);
  for (int i=0;i<m->n_vox*s->d.n_upper;i++) s->S[i]=(REAL) S[i];
}

/* RT_grid::brightness (RT_grid.hpp:233-299) with emission_voxels::update_tracker_brightness_{interp,nointerp}
   (:199-233), multiplet update_tracker_start_interp (:352-378) and update_tracker_brightness (:302-316).
   out[(3 n_lines + n_lower)][n_los]: brightness[line], tau_species_final[line], tau_absorber_final[line], col[lower] */
void oracle_multiplet_brightness(void *h, int n_los, const double *loc, const double *dir, int n_subsamples, double *out) {
  omodel *m=(omodel*) h; omstate *s=om_get(m); const omult *d=&s->d;
  const int NL=d->n_lines, NLOW=d->n_lower, NUP=d->n_upper;
#pragma omp parallel
  {
    obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
    omtracker *t=(omtracker*) malloc(sizeof(omtracker));
#pragma omp for schedule(dynamic,64)
    for (int i=0;i<n_los;i++) {
      ovec v; o_los(loc+3*i, dir+3*i, &v);
      int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
      om_reset(d,t,0.0,NULL);
      if (n>0) {
	int nsd = n_subsamples; if (n_subsamples==0) nsd=2;
	for (int ib=1;ib<n;ib++) {
	  REAL d_start=b[begin+ib-1].distance;
	  REAL d_step=(b[begin+ib].distance-d_start)/(nsd-1);
	  const REAL eps=O_EPS;
	  d_start += RL(0.5)*eps*d_step;
	  d_step *= RL(1.0)-eps;
	  int cur=b[begin+ib-1].entering;
	  for (int is=1;is<nsd;is++) {
	    REAL pr, pt; o_extend(&v, d_start+is*d_step, &pr, &pt);
	    REAL T, dens[M_MAXLOW], nabs, Sv[M_MAXUP];
	    if (n_subsamples==0) {
	      T=s->T[cur]; nabs=s->nabs[cur];
	      for (int l=0;l<NLOW;l++) dens[l]=s->n[l][cur];
	      for (int u=0;u<NUP;u++) Sv[u]=s->S[cur*NUP+u];
	    } else {
	      int idx[4]; REAL w[4];
	      o_interp_weights(m,cur,pr,pt,idx,w);
	      T=o_interp(s->T_pt,idx,w);
	      for (int l=0;l<NLOW;l++) dens[l]=o_interp(s->n_pt[l],idx,w);
	      nabs=o_interp(s->nabs_pt,idx,w);
	      for (int u=0;u<NUP;u++) { REAL a=0; for (int k=0;k<4;k++) a += w[k]*s->S[idx[k]*NUP+u]; Sv[u]=a; }
	    }
	    om_update(d,0,T,dens,nabs,d_step,t);
	    for (int line=0;line<NL;line++)
	      t->brightness[line] += (Sv[d->upper[line]]*d->A[line]*t->T_int[line]/RL(1e9));
	  }
	}
	if (eb) for (int line=0;line<NL;line++) t->tau_ab[line]=-1.0;
      }
      for (int line=0;line<NL;line++) {
	out[(size_t)(0*NL+line)*n_los+i]=t->brightness[line];
	out[(size_t)(1*NL+line)*n_los+i]=t->tau_sp[line];
	out[(size_t)(2*NL+line)*n_los+i]=t->tau_ab[line];
      }
      for (int l=0;l<NLOW;l++) out[(size_t)(3*NL+l)*n_los+i]=t->col[l];
    }
    free(b); free(t);
  }
}
