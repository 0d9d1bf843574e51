/* rt_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of planetarymike/3D_planetary_RT_model's influence-matrix
 * build, source-function solve and line-of-sight brightness for singlet CFR
 * emissions on the spherical azimuthally-symmetric grid.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/src).
 *
 * PINNING: tests/test_oracle_vs_ref.py checks this file against the reference's
 * own source compiled in place (oracle/_ref, see oracle/Makefile) -- boundary
 * lists bit for bit, K / S0 / S / brightness to rounding -- and tests/golden/
 * holds fixtures produced by that reference build for the GPU box, where
 * /root/reference does not exist.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library; the product (lib3d_planetary_rt_b200.so) never does.
 *
 * Precision: REAL = double, or float with -DORACLE_FLOAT (reference Real.hpp:9-27).
 * Where the reference calls an unqualified libm function on a Real it gets the
 * double version even in float builds (global ::cos etc.); where it calls std::cos
 * it gets the float overload.  The STD_* macros reproduce that.
 * Arithmetic must be IEEE without FMA contraction: build with -ffp-contract=off.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef ORACLE_FLOAT
typedef float REAL;
#define RL(x) (x##f)
#define O_EPS 1e-3f
#define O_STRICTEPS 1e-5f
#define O_CONEEPS 1e-2f
#define STD_COS cosf
#define STD_SIN sinf
#define STD_SQRT sqrtf
#define STD_EXP expf
#define STD_LOG logf
#define STD_ABS fabsf
#else
typedef double REAL;
#define RL(x) (x)
#define O_EPS 1e-6
#define O_STRICTEPS 1e-10
#define O_CONEEPS O_EPS
#define STD_COS cos
#define STD_SIN sin
#define STD_SQRT sqrt
#define STD_EXP exp
#define STD_LOG log
#define STD_ABS fabs
#endif

/* constants.hpp:28-30 */
static const REAL o_pi = (REAL) M_PI;
static const REAL o_two_over_sqrt_pi = (REAL) M_2_SQRTPI;
static const REAL o_one_over_sqrt_pi = (REAL)(((REAL) M_2_SQRTPI)/2.0);
#define O_ONE_OVER_SQRT_PI o_one_over_sqrt_pi

#define N_LAMBDA 20               /* los_tracker.hpp:122 */
static const REAL lambda_max = RL(4.0);  /* los_tracker.hpp:123 */

typedef struct {
  int n_rb, n_sb, n_theta, n_phi, n_vox, n_rays, cap;
  int pp;   /* 1 = plane_parallel_grid (grid/grid_plane_parallel.hpp): n_sb = 2, planes z = rb[i] */
  REAL rmin, rmax;
  REAL *rb, *pts_r, *log_pts_r, *sph_R, *sph_R2;        /* radial */
  REAL *sb, *pts_s, *cone_cos, *cone_cos2;              /* sza    */
  REAL *vx, *vy, *vz, *vr, *vt;                         /* voxel points (p = 0) */
  REAL *ray_t, *ray_p, *ray_cost, *ray_sint, *ray_domega;
  /* emissions */
  int n_em;
  REAL branching[2], T_ref[2], sigma_ref[2], g_factor[2];
  REAL *arr[2][10];   /* T_ratio, T_ratio_pt, density, density_pt, dtau_sp, dtau_sp_pt, dtau_abs, dtau_abs_pt, abs, abs_pt */
  REAL *K[2], *S0[2], *tau_sp_ss[2], *tau_abs_ss[2], *S[2];
  void *mult_state;   /* multiplet emission (multiplet_oracle.inc.c), allocated on first use */
} omodel;

enum { A_TR=0, A_TR_PT, A_N, A_N_PT, A_DTS, A_DTS_PT, A_DTA, A_DTA_PT, A_ABS, A_ABS_PT };

static REAL *ralloc(size_t n) { return (REAL*) calloc(n ? n : 1, sizeof(REAL)); }

/* gauss_legendre_quadrature.cpp:9-46 */
static void o_gauleg(REAL x1, REAL x2, REAL *x, REAL *w, int n) {
  REAL z1, z, xm, xl, pp, p3, p2, p1;
  int m = (n+1)/2;
  xm = 0.5*(x2+x1);
  xl = 0.5*(x2-x1);
  for (int i=0;i<m;i++) {
    z = cos(M_PI*(i+0.75)/(n+0.5));
    do {
      p1 = 1.0; p2 = 0.0;
      for (int j=0;j<n;j++) { p3=p2; p2=p1; p1=((2*j+1)*z*p2 - j*p3)/(j+1); }
      pp = n*(z*p1-p2)/(z*z-1.0);
      z1 = z;
      z = z1 - p1/pp;
    } while (STD_ABS(z-z1) > O_STRICTEPS);
    x[i] = xm - xl*z;
    x[n-1-i] = xm + xl*z;
    w[i] = 2.0*xl/((1.0-z*z)*pp*pp);
    w[n-1-i] = w[i];
  }
}

/* grid_spherical_azimuthally_symmetric.hpp:159-406 (setup_voxels from given radial
   boundaries, setup_rays); intersections.cpp:53-56,103-109; atmo_vec.cpp:41-49,172-190 */
void* oracle_create(int n_rb, int n_sb, int n_theta, int n_phi,
		    const double *rb_in, int szamethod, int raymethod) {
  omodel *m = (omodel*) calloc(1, sizeof(omodel));
  m->n_rb=n_rb; m->n_sb=n_sb; m->n_theta=n_theta; m->n_phi=n_phi;
  m->n_vox=(n_rb-1)*(n_sb-1); m->n_rays=n_theta*n_phi; m->cap=2*n_rb+n_sb;
  m->rb=ralloc(n_rb); m->pts_r=ralloc(n_rb-1); m->log_pts_r=ralloc(n_rb-1);
  m->sph_R=ralloc(n_rb); m->sph_R2=ralloc(n_rb);
  m->sb=ralloc(n_sb); m->pts_s=ralloc(n_sb-1); m->cone_cos=ralloc(n_sb-2); m->cone_cos2=ralloc(n_sb-2);
  m->vx=ralloc(m->n_vox); m->vy=ralloc(m->n_vox); m->vz=ralloc(m->n_vox); m->vr=ralloc(m->n_vox); m->vt=ralloc(m->n_vox);
  m->ray_t=ralloc(m->n_rays); m->ray_p=ralloc(m->n_rays); m->ray_cost=ralloc(m->n_rays);
  m->ray_sint=ralloc(m->n_rays); m->ray_domega=ralloc(m->n_rays);

  for (int i=0;i<n_rb;i++) m->rb[i]=(REAL) rb_in[i];
  m->rmin=(REAL) rb_in[0]; m->rmax=(REAL) rb_in[n_rb-1];
  for (int i=0;i<n_rb-1;i++) {
    m->pts_r[i]=sqrt(m->rb[i]*m->rb[i+1]);           /* :303 (unqualified sqrt) */
    m->log_pts_r[i]=STD_LOG(m->pts_r[i]);             /* :304 (std::log via using-decl) */
  }
  const REAL scale = 1e9;                              /* intersections.hpp:12 */
  for (int i=0;i<n_rb;i++) { m->sph_R[i]=m->rb[i]/scale; m->sph_R2[i]=m->sph_R[i]*m->sph_R[i]; }

  if (szamethod==0) {                                  /* szamethod_uniform :314-319 */
    REAL sza_spacing = o_pi/(n_sb-2.);
    for (int i=0;i<n_sb;i++) m->sb[i]=(i-0.5)*sza_spacing;
  } else {                                             /* szamethod_uniform_cos :320-329 */
    REAL cs = 2.0/(n_sb-2.);
    m->sb[0] = -acos(1.0-0.5*cs);
    for (int i=1;i<n_sb-1;i++) m->sb[i]=acos(1.0-(i-0.5)*cs);
    m->sb[n_sb-1] = o_pi + acos(1.0-0.5*cs);
  }
  for (int i=0;i<n_sb-1;i++) m->pts_s[i]=0.5*(m->sb[i]+m->sb[i+1]);          /* :331-333 */
  for (int i=0;i<n_sb-2;i++) { m->cone_cos[i]=STD_COS(m->sb[i+1]); m->cone_cos2[i]=m->cone_cos[i]*m->cone_cos[i]; }

  for (int i=0;i<n_rb-1;i++)
    for (int j=0;j<n_sb-1;j++) {                       /* :344-361, atmo_point::rtp */
      int v=i*(n_sb-1)+j;
      REAL r=m->pts_r[i], t=m->pts_s[j], p=0.;
      m->vr[v]=r; m->vt[v]=t;
      m->vx[v]=r*sin(t)*cos(p);
      m->vy[v]=r*sin(t)*sin(p);
      m->vz[v]=r*cos(t);
    }

  /* rays :365-406 */
  REAL *th=ralloc(n_theta), *wt=ralloc(n_theta);
  if (raymethod==0) {
    o_gauleg(0, o_pi, th, wt, n_theta);
    for (int i=0;i<n_theta;i++) wt[i]*=STD_SIN(th[i]);
  } else {
    REAL theta_spacing = o_pi/(n_theta-1);
    for (int i=0;i<n_theta;i++) {
      th[i]=i*theta_spacing;
      if (i==0 || i==n_theta-1) wt[i]=1-STD_COS(theta_spacing/2);
      else wt[i]=(STD_COS(th[i]-theta_spacing/2)-STD_COS(th[i]+theta_spacing/2));
    }
  }
  REAL phi_spacing = 2*o_pi/n_phi;
  for (int i=0;i<n_theta;i++)
    for (int j=0;j<n_phi;j++) {
      int k=i*n_phi+j;
      REAL ph=(j+0.5)*phi_spacing;
      m->ray_t[k]=th[i]; m->ray_p[k]=ph;
      m->ray_cost[k]=STD_COS(th[i]); m->ray_sint[k]=STD_SIN(th[i]);   /* atmo_ray::tp */
      m->ray_domega[k]=wt[i]*phi_spacing*RL(0.25)/o_pi;               /* set_ray_index */
    }
  free(th); free(wt);
  return m;
}

/* plane_parallel_grid<NR, NTH>::setup_voxels / setup_rays (grid_plane_parallel.hpp:189-223) from given
   radial boundaries: voxel i has the point xyz(0,0,pts_radii[i]); rays = Gauss-Legendre in theta on
   [0, pi] with weights w*sin(theta), phi = 0, domega = w*2pi/(4pi) */
void* oracle_create_pp(int n_rb, int n_theta, const double *rb_in) {
  omodel *m = (omodel*) calloc(1, sizeof(omodel));
  m->pp=1;
  m->n_rb=n_rb; m->n_sb=2; m->n_theta=n_theta; m->n_phi=1;
  m->n_vox=n_rb-1; m->n_rays=n_theta; m->cap=n_rb+1;          /* n_max_intersections = n_rb (+ origin slack) */
  m->rb=ralloc(n_rb); m->pts_r=ralloc(n_rb-1); m->log_pts_r=ralloc(n_rb-1);
  m->sph_R=ralloc(n_rb); m->sph_R2=ralloc(n_rb);
  m->sb=ralloc(2); m->pts_s=ralloc(1); m->cone_cos=ralloc(1); m->cone_cos2=ralloc(1);
  m->vx=ralloc(m->n_vox); m->vy=ralloc(m->n_vox); m->vz=ralloc(m->n_vox); m->vr=ralloc(m->n_vox); m->vt=ralloc(m->n_vox);
  m->ray_t=ralloc(m->n_rays); m->ray_p=ralloc(m->n_rays); m->ray_cost=ralloc(m->n_rays);
  m->ray_sint=ralloc(m->n_rays); m->ray_domega=ralloc(m->n_rays);
  for (int i=0;i<n_rb;i++) m->rb[i]=(REAL) rb_in[i];
  m->rmin=(REAL) rb_in[0]; m->rmax=(REAL) rb_in[n_rb-1];
  m->sb[0]=0; m->sb[1]=o_pi;
  for (int i=0;i<n_rb-1;i++) {
    m->pts_r[i]=sqrt(m->rb[i]*m->rb[i+1]);                      /* :198 */
    m->log_pts_r[i]=STD_LOG(m->pts_r[i]);
    /* atmo_point::xyz(0,0,z) atmo_vec.cpp:51-61 */
    REAL z=m->pts_r[i];
    m->vx[i]=0.; m->vy[i]=0.; m->vz[i]=z;
    m->vr[i]=hypot(hypot((REAL)0.,(REAL)0.),z);
    m->vt[i]=acos(z/m->vr[i]);
  }
  REAL *th=ralloc(n_theta), *wt=ralloc(n_theta);
  o_gauleg(0, o_pi, th, wt, n_theta);
  for (int i=0;i<n_theta;i++) wt[i]*=STD_SIN(th[i]);
  for (int i=0;i<n_theta;i++) {
    m->ray_t[i]=th[i]; m->ray_p[i]=0.0;
    m->ray_cost[i]=STD_COS(th[i]); m->ray_sint[i]=STD_SIN(th[i]);
    m->ray_domega[i]=wt[i]*(2*o_pi)*RL(0.25)/o_pi;
  }
  free(th); free(wt);
  return m;
}

static void om_free(omodel *m);
void oracle_destroy(void *h) {
  omodel *m=(omodel*) h;
  if (!m) return;
  REAL *ps[] = {m->rb,m->pts_r,m->log_pts_r,m->sph_R,m->sph_R2,m->sb,m->pts_s,m->cone_cos,m->cone_cos2,
		m->vx,m->vy,m->vz,m->vr,m->vt,m->ray_t,m->ray_p,m->ray_cost,m->ray_sint,m->ray_domega};
  for (size_t i=0;i<sizeof(ps)/sizeof(ps[0]);i++) free(ps[i]);
  for (int e=0;e<2;e++) {
    for (int a=0;a<10;a++) free(m->arr[e][a]);
    free(m->K[e]); free(m->S0[e]); free(m->tau_sp_ss[e]); free(m->tau_abs_ss[e]); free(m->S[e]);
  }
  om_free(m);
  free(m);
}

int oracle_real_bytes(void) { return (int) sizeof(REAL); }

void oracle_get_grid(void *h, double *sza_b, double *pts_r, double *pts_sza, double *ray_theta,
		     double *ray_phi, double *ray_domega) {
  omodel *m=(omodel*) h;
  for (int i=0;i<m->n_sb;i++) sza_b[i]=m->sb[i];
  for (int i=0;i<m->n_rb-1;i++) pts_r[i]=m->pts_r[i];
  for (int i=0;i<m->n_sb-1;i++) pts_sza[i]=m->pts_s[i];
  for (int i=0;i<m->n_theta;i++) ray_theta[i]=m->ray_t[i*m->n_phi];
  for (int j=0;j<m->n_phi;j++) ray_phi[j]=m->ray_p[j];
  for (int k=0;k<m->n_rays;k++) ray_domega[k]=m->ray_domega[k];
}

/* singlet_CFR::define, singlet_CFR.hpp:419-492.  vox_in = [6][n_vox]:
   n_avg, n_pt, T_avg, T_pt, nabs_avg, nabs_pt (what the atmosphere's *_voxel_avg
   callbacks return); abs_sigma = absorber cross section (T-independent) */
void oracle_define_singlet(void *h, int e, double branching, double T_ref, double sigma_ref,
			   double g_factor, double abs_sigma, const double *vox_in) {
  omodel *m=(omodel*) h;
  const int N=m->n_vox;
  if (e+1>m->n_em) m->n_em=e+1;
  m->branching[e]=(REAL) branching; m->T_ref[e]=(REAL) T_ref;
  m->sigma_ref[e]=(REAL) sigma_ref; m->g_factor[e]=(REAL) g_factor;
  for (int a=0;a<10;a++) { free(m->arr[e][a]); m->arr[e][a]=ralloc(N); }
  free(m->K[e]); m->K[e]=ralloc((size_t) N*N);
  free(m->S0[e]); m->S0[e]=ralloc(N); free(m->S[e]); m->S[e]=ralloc(N);
  free(m->tau_sp_ss[e]); m->tau_sp_ss[e]=ralloc(N); free(m->tau_abs_ss[e]); m->tau_abs_ss[e]=ralloc(N);
  for (int i=0;i<N;i++) {
    REAL n_avg=(REAL) vox_in[0*N+i], n_pt=(REAL) vox_in[1*N+i];
    REAL T_avg=(REAL) vox_in[2*N+i], T_pt=(REAL) vox_in[3*N+i];
    REAL a_avg=(REAL) vox_in[4*N+i], a_pt=(REAL) vox_in[5*N+i];
    REAL sig=(REAL) abs_sigma;
    m->arr[e][A_N][i]=n_avg; m->arr[e][A_N_PT][i]=n_pt;
    m->arr[e][A_TR][i]=m->T_ref[e]/T_avg;                    /* :451-452 */
    m->arr[e][A_TR_PT][i]=m->T_ref[e]/T_pt;
    /* :482-487, coefficient-wise, evaluated left to right */
    m->arr[e][A_DTS][i]   =n_avg*m->sigma_ref[e]*STD_SQRT(m->arr[e][A_TR][i]);
    m->arr[e][A_DTS_PT][i]=n_pt *m->sigma_ref[e]*STD_SQRT(m->arr[e][A_TR_PT][i]);
    m->arr[e][A_DTA][i]   =a_avg*sig;
    m->arr[e][A_DTA_PT][i]=a_pt*sig;
    m->arr[e][A_ABS][i]   =m->arr[e][A_DTA][i]/m->arr[e][A_DTS][i];
    m->arr[e][A_ABS_PT][i]=m->arr[e][A_DTA_PT][i]/m->arr[e][A_DTS_PT][i];
  }
}

void oracle_get_arrays(void *h, int e, double *out) {
  omodel *m=(omodel*) h;
  for (int a=0;a<10;a++) for (int i=0;i<m->n_vox;i++) out[(size_t)a*m->n_vox+i]=m->arr[e][a][i];
}

/* ------------------------------------------------------------------ traversal */
typedef struct { int entering; int idx[2]; REAL distance; } obnd;   /* boundaries.hpp:15-61 */

typedef struct {  /* the part of atmo_vector the traversal / brightness use */
  REAL x,y,z,r,t; int i_voxel;
  REAL lx,ly,lz,cost;
} ovec;

static int o_samesign(REAL a, REAL b) {               /* intersections.cpp:7-12 */
  return ((a>0&&b>0) || (a<0&&b<0) || (a==0&&b==0));
}
static int o_is_zero(REAL a, REAL tol) { return !(a>tol || a<-tol); }   /* :14-19 */

/* sphere::intersections, intersections.cpp:58-95 */
static void o_sphere(const omodel *m, int ir, const ovec *v, REAL *d, int *nh) {
  const REAL scale = 1e9;
  *nh=0;
  const REAL r_norm = v->r/scale;
  const REAL B = r_norm*v->cost;
  const REAL C = r_norm*r_norm - m->sph_R2[ir];
  const REAL discr = B*B-C;
  if (discr > 0) {
    const REAL d0 = (B>0) ? -B-STD_SQRT(discr) : -B+STD_SQRT(discr);
    if (d0>0) { d[*nh]=d0*scale; (*nh)++; }
    const REAL d1 = C/d0;
    if (d1>0) { d[*nh]=d1*scale; (*nh)++; }
  }
}

/* cone::intersections, intersections.cpp:120-179 */
static void o_cone(const omodel *m, int k, const ovec *v, REAL *d, int *nh) {
  *nh=0;
  const REAL cosangle=m->cone_cos[k], cosangle2=m->cone_cos2[k];
  const REAL rscale = v->r;
  const REAL z_norm = v->z/rscale;
  const REAL A = v->lz*v->lz - cosangle2;
  const REAL B = z_norm*v->lz - v->cost*cosangle2;
  const REAL C = z_norm*z_norm - cosangle2;
  if (!o_is_zero(A, O_STRICTEPS)) {
    const REAL discr = B*B-A*C;
    if (discr > 0) {
      const REAL q = (B>0) ? -B-STD_SQRT(discr) : -B+STD_SQRT(discr);
      const REAL d0 = q/A;
      if (d0>0 && o_samesign(z_norm+d0*v->lz, cosangle)) { d[*nh]=d0*rscale; (*nh)++; }
      const REAL d1 = C/q;
      if (d1>0 && o_samesign(z_norm+d1*v->lz, cosangle)) { d[*nh]=d1*rscale; (*nh)++; }
    }
  } else {
    const REAL dd = -C/(2*B);
    if (dd>0 && o_samesign(z_norm+dd*v->lz, cosangle)) { d[*nh]=dd*rscale; (*nh)++; }
  }
}

/* boundary_set::add_intersections, boundaries.hpp:160-193 */
static void o_add(obnd *b, int *n, REAL start, int dim, int idx, REAL coord, const REAL *d, int nh) {
  if (nh==0) return;
  int above = (start > coord);
  obnd nb; nb.entering=-2; nb.idx[0]=nb.idx[1]=-2; nb.distance=-1;
  if (nh==1) {
    nb.idx[dim] = above ? idx-1 : idx;
    nb.distance = d[0];
    b[(*n)++]=nb;
  } else {
    int in_order = (d[1] > d[0]);
    nb.idx[dim] = above ? idx-1 : idx;
    nb.distance = in_order ? d[0] : d[1];
    b[(*n)++]=nb;
    nb.idx[dim] = above ? idx : idx-1;
    nb.distance = in_order ? d[1] : d[0];
    b[(*n)++]=nb;
  }
}

static int o_vox(const omodel *m, int ri, int si) {   /* indices_to_voxel :408-415 */
  if (ri<0 || ri>m->n_rb-2 || si<0 || si>m->n_sb-2) return -1;
  return ri*(m->n_sb-1)+si;
}
static int o_find(REAL c, const REAL *bnd, int n) {   /* find_coordinate_index :433-450 */
  int i;
  for (i=0;i<n;i++) if (c < bnd[i]) break;
  return i-1;
}

/* spherical_azimuthally_symmetric_grid::ray_voxel_intersections :459-509 with
   boundary_set::{sort,propagate_indices,assign_voxel_indices,trim} boundaries.hpp:131-232.
   b must hold cap = 2*n_rb+n_sb entries.  Returns the trimmed length; *begin = offset
   of the first kept entry; *exits_bottom as boundary_intersection_stepper::init_stepper :334-349 */
static int o_traverse(const omodel *m, const ovec *v, obnd *b, int *begin_out, int *exits_bottom) {
  int n=0;
  obnd o; o.distance=0.0;
  if (v->i_voxel==-1) {
    o.idx[0]=o_find(v->r, m->rb, m->n_rb);
    o.idx[1]=o_find(v->t, m->sb, m->n_sb);
    o.entering=o_vox(m,o.idx[0],o.idx[1]);
  } else {
    o.entering=v->i_voxel;
    if (v->i_voxel<0 || v->i_voxel>m->n_vox-1) { o.idx[0]=o.idx[1]=-1; }
    else { o.idx[0]=v->i_voxel/(m->n_sb-1); o.idx[1]=v->i_voxel%(m->n_sb-1); }
  }
  if (m->pp) {   /* plane_parallel_grid::point_to_indices / voxel_to_indices: one dimension */
    if (v->i_voxel==-1) { o.idx[0]=o_find(v->r, m->rb, m->n_rb); o.idx[1]=0; o.entering=o_vox(m,o.idx[0],0); }
    else { o.idx[1]=0; }
  }
  b[n++]=o;
  REAL d[2]={-1,-1}; int nh=0;
  if (m->pp) {   /* plane::intersections intersections.cpp:25-46, grid_plane_parallel.hpp:282-288 */
    for (int ir=0;ir<m->n_rb;ir++) {
      nh=0;
      if (v->lz != 0) {
	REAL dd=(m->rb[ir]-v->z)/v->lz;
	if (dd > 0) { d[nh]=dd; nh++; }
      }
      o_add(b,&n,v->r,0,ir,m->rb[ir],d,nh);
    }
  } else
  for (int ir=0;ir<m->n_rb;ir++) {
    o_sphere(m,ir,v,d,&nh);
    o_add(b,&n,v->r,0,ir,m->rb[ir],d,nh);
  }
  for (int k=0;k<m->n_sb-2;k++) {
    o_cone(m,k,v,d,&nh);
    o_add(b,&n,v->t,1,k+1,m->sb[k+1],d,nh);
  }
  /* insertion sort, strict <  (stable) */
  for (int i=1;i<n;i++) {
    obnd key=b[i]; int j=i-1;
    while (j>=0 && key.distance < b[j].distance) { b[j+1]=b[j]; j--; }
    b[j+1]=key;
  }
  for (int i=1;i<n;i++)
    for (int j=0;j<2;j++) if (b[i].idx[j]==-2) b[i].idx[j]=b[i-1].idx[j];
  for (int i=1;i<n;i++) b[i].entering=o_vox(m,b[i].idx[0],b[i].idx[1]);
  /* trim */
  int begin=0, size=n;
  while (b[begin].entering==-1 && begin<n-1) begin++;
  if (begin==n-1) { begin=0; size=0; }
  else {
    int end=begin;
    do { end++; } while (b[end].entering!=-1 && end<n-1);
    size=end-begin+1;
  }
  *begin_out=begin;
  *exits_bottom = (size>0 && b[begin+size-1].idx[0]==-1) ? 1 : 0;
  return size;
}

/* atmo_vector::ptray, atmo_vec.cpp:228-249 (voxel points have p = 0) */
static void o_ptray(const omodel *m, int iv, int ir, ovec *v) {
  v->x=m->vx[iv]; v->y=m->vy[iv]; v->z=m->vz[iv]; v->r=m->vr[iv]; v->t=m->vt[iv]; v->i_voxel=iv;
  const REAL ptp=0., ptt=m->vt[iv], rayp=m->ray_p[ir], cost=m->ray_cost[ir], sint=m->ray_sint[ir];
  v->cost=cost;
  v->lx = (cos(ptp)*cos(rayp)*cos(ptt)*sint + cos(ptp)*cost*sin(ptt) - sint*sin(rayp)*sin(ptp));
  v->ly = (cost*sin(ptp)*sin(ptt) + sint*cos(rayp)*sin(ptp)*cos(ptt) + sint*sin(rayp)*cos(ptp));
  v->lz = cost*cos(ptt)-cos(rayp)*sint*sin(ptt);
}

/* atmo_point::xyz atmo_vec.cpp:51-61 + atmo_vector::ptxyz :256-290 */
static void o_ptxyz(REAL x, REAL y, REAL z, int i_voxel_known, REAL r_known, REAL t_known,
		    REAL dx, REAL dy, REAL dz, ovec *v) {
  v->x=x; v->y=y; v->z=z;
  if (i_voxel_known>=0) { v->r=r_known; v->t=t_known; v->i_voxel=i_voxel_known; }
  else { v->r=hypot(hypot(x,y),z); v->t=acos(z/v->r); v->i_voxel=-1; }
  REAL mag=hypot(hypot(dx,dy),dz);
  v->lx=dx/mag; v->ly=dy/mag; v->lz=dz/mag;
  REAL costx=v->lx*(v->x/v->r);
  REAL costy=v->ly*(v->y/v->r);
  REAL costz=v->lz*(v->z/v->r);
  v->cost=costx+costy+costz;
}

/* observation::add_MSO_observation observation.hpp:46-65: model = (MSO_z, -MSO_y, MSO_x) */
static void o_los(const double *loc, const double *dir, ovec *v) {
  REAL l0=(REAL) loc[0], l1=(REAL) loc[1], l2=(REAL) loc[2];
  REAL d0=(REAL) dir[0], d1=(REAL) dir[1], d2=(REAL) dir[2];
  o_ptxyz(l2, -l1, l0, -1, 0, 0, d2, -d1, d0, v);
}

static long o_dump(const obnd *b, int begin, int len, int eb, long pos, long cap, int *len_out, int *eb_out,
		   int *entering, double *distance) {
  *len_out=len; *eb_out=eb;
  if (pos+len>cap) return -1;
  for (int k=0;k<len;k++) { entering[pos+k]=b[begin+k].entering; distance[pos+k]=b[begin+k].distance; }
  return pos+len;
}

long oracle_traverse_voxel_rays(void *h, int v0, int v1, long cap, int *len, int *exits_bottom,
				int *entering, double *distance) {
  omodel *m=(omodel*) h;
  obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
  long pos=0;
  for (int iv=v0;iv<v1;iv++)
    for (int ir=0;ir<m->n_rays;ir++) {
      ovec v; o_ptray(m,iv,ir,&v);
      int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
      long idx=(long)(iv-v0)*m->n_rays+ir;
      pos=o_dump(b,begin,n,eb,pos,cap,len+idx,exits_bottom+idx,entering,distance);
      if (pos<0) { free(b); return -1; }
    }
  free(b);
  return pos;
}

long oracle_traverse_los(void *h, int n_los, const double *loc, const double *dir, long cap, int *len,
			 int *exits_bottom, int *entering, double *distance, double *rayscal) {
  omodel *m=(omodel*) h;
  obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
  long pos=0;
  for (int i=0;i<n_los;i++) {
    ovec v; o_los(loc+3*i, dir+3*i, &v);
    if (rayscal) { rayscal[6*i+0]=v.r; rayscal[6*i+1]=v.z; rayscal[6*i+2]=v.t; rayscal[6*i+3]=v.cost; rayscal[6*i+4]=v.lz; rayscal[6*i+5]=v.lx; }
    int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
    pos=o_dump(b,begin,n,eb,pos,cap,len+i,exits_bottom+i,entering,distance);
    if (pos<0) { free(b); return -1; }
  }
  free(b);
  return pos;
}

/* ------------------------------------------------------------------ singlet CFR physics */
typedef struct {   /* los_tracker.hpp:11-170 */
  REAL tau_species_final, tau_absorber_final, species_col_dens;
  REAL holstein_T_final, holstein_T_int, holstein_G_int, brightness;
  REAL T_ratio_at_origin;
  REAL P[N_LAMBDA];
} otracker;

static void o_reset(otracker *t, REAL T_ratio) {   /* los_tracker.hpp:41-50,161-168 */
  t->tau_species_final=0.0; t->tau_absorber_final=0.0; t->holstein_T_final=1.0; t->species_col_dens=0.0;
  t->brightness=0.0;
  t->T_ratio_at_origin=T_ratio;
  for (int i=0;i<N_LAMBDA;i++) t->P[i]=1.0;
}

static REAL o_lineshape(int i, REAL T_ratio) {     /* los_tracker.hpp:127-147 */
  const REAL delta_lambda = lambda_max/(N_LAMBDA-1);
  REAL lambda2 = i*delta_lambda;
  lambda2 *= lambda2;
  return STD_EXP(-lambda2*T_ratio);
}
static REAL o_weight(int i) {                      /* los_tracker.hpp:131-136 */
  const REAL delta_lambda = lambda_max/(N_LAMBDA-1);
  if (i==0 || i==N_LAMBDA-1) return delta_lambda;
  return RL(2.0)*delta_lambda;
}
static REAL o_norm(REAL T_ratio) {                 /* los_tracker.hpp:150-154 */
  return O_ONE_OVER_SQRT_PI*STD_SQRT(T_ratio);
}

/* singlet_CFR::update_tracker_start<influence>, singlet_CFR.hpp:80-260 */
static void o_update(int influence, REAL Tr, REAL dens, REAL dts, REAL dta, REAL pathlength, otracker *t) {
  REAL col = dens*pathlength;
  t->species_col_dens += col;
  REAL tau_species_voxel = dts*pathlength;
  t->tau_species_final += tau_species_voxel;
  t->tau_absorber_final += dta*pathlength;
  t->holstein_T_int=0; t->holstein_T_final=0; t->holstein_G_int=0;
  for (int i=0;i<N_LAMBDA;i++) {
    REAL lineshape = o_lineshape(i, Tr);
    REAL tau_lambda_voxel = ((dta + (dts*lineshape))*pathlength);
    REAL tp_voxel = STD_EXP(-tau_lambda_voxel);
    REAL tp_final = (t->P[i]*tp_voxel);
    REAL holTcoef = o_weight(i);
    REAL coef;
    if (tau_lambda_voxel < 1e-3) coef = (RL(1.0)-(RL(0.5)*tau_lambda_voxel));
    else coef = ((RL(1.0)-tp_voxel)/(tau_lambda_voxel));
    coef *= (holTcoef*lineshape*t->P[i]*dts*pathlength);
    t->holstein_T_int += coef;
    if (influence) {
      REAL ls0 = o_lineshape(i, t->T_ratio_at_origin);
      REAL renorm = o_norm(t->T_ratio_at_origin);
      t->holstein_T_final += (holTcoef*renorm*ls0*tp_final);
      t->holstein_G_int += (coef*renorm*ls0);
    }
    t->P[i]=tp_final;
  }
  if (t->holstein_T_int > tau_species_voxel) t->holstein_T_int = tau_species_voxel;
}

/* RT_grid::generate_S loop body, RT_grid.hpp:166-201: rows v0..v1-1 (stride) of K, plus S0
   and the single-scattering optical depths; accumulate_influence emission_voxels.hpp:137-155.
   Returns the number of ray-voxel steps (one step = one influence_update, all emissions). */
long oracle_build_rows(void *h, int v0, int v1, int stride) {
  omodel *m=(omodel*) h;
  const int N=m->n_vox;
  long steps=0;
#pragma omp parallel reduction(+:steps)
  {
    obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
    REAL *infl=ralloc((size_t) N*m->n_em);
    int *touched=(int*) malloc(sizeof(int)*m->cap);
#pragma omp for schedule(dynamic,1)
    for (int iv=v0;iv<v1;iv+=stride) {
      for (int e=0;e<m->n_em;e++) memset(m->K[e]+(size_t)iv*N, 0, sizeof(REAL)*N);
      for (int ir=0;ir<m->n_rays;ir++) {
	ovec v; o_ptray(m,iv,ir,&v);
	int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
	if (n==0) continue;
	otracker t[2];
	for (int e=0;e<m->n_em;e++) o_reset(&t[e], m->arr[e][A_TR][iv]);
	int nt=0;
	for (int k=1;k<n;k++) {                                /* boundaries.hpp:358-377 */
	  int vox=b[begin+k-1].entering;
	  REAL s=b[begin+k].distance-b[begin+k-1].distance;
	  touched[nt++]=vox;
	  for (int e=0;e<m->n_em;e++) {                        /* singlet_CFR.hpp:352-370 */
	    REAL **a=m->arr[e];
	    o_update(1, a[A_TR][vox], a[A_N][vox], a[A_DTS][vox], a[A_DTA][vox], s, &t[e]);
	    REAL coef=m->ray_domega[ir];
	    coef*=t[e].holstein_G_int;
	    infl[(size_t)e*N+vox]+=coef;
	  }
	  steps++;
	}
	for (int e=0;e<m->n_em;e++) {
	  REAL *row=m->K[e]+(size_t)iv*N;
	  for (int q=0;q<nt;q++) { int vox=touched[q]; REAL c=infl[(size_t)e*N+vox]; if (c!=0) { row[vox]+=c; infl[(size_t)e*N+vox]=0; } }
	}
      }
      /* single scattering, RT_grid.hpp:121-139; singlet_CFR.hpp:372-398 */
      if (m->vz[iv]<0 && m->vx[iv]*m->vx[iv]+m->vy[iv]*m->vy[iv] < m->rmin*m->rmin) {
	for (int e=0;e<m->n_em;e++) { m->tau_sp_ss[e][iv]=RL(-1.0); m->tau_abs_ss[e][iv]=RL(-1.0); m->S0[e][iv]=RL(0.0); }
      } else {
	ovec v; o_ptxyz(m->vx[iv],m->vy[iv],m->vz[iv], iv, m->vr[iv], m->vt[iv], 0., 0., 1., &v);
	int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
	otracker t[2];
	for (int e=0;e<m->n_em;e++) o_reset(&t[e], m->arr[e][A_TR][iv]);
	for (int k=1;k<n;k++) {
	  int vox=b[begin+k-1].entering;
	  REAL s=b[begin+k].distance-b[begin+k-1].distance;
	  for (int e=0;e<m->n_em;e++) {
	    REAL **a=m->arr[e];
	    o_update(1, a[A_TR][vox], a[A_N][vox], a[A_DTS][vox], a[A_DTA][vox], s, &t[e]);
	  }
	}
	for (int e=0;e<m->n_em;e++) {
	  m->tau_sp_ss[e][iv]=t[e].tau_species_final; m->tau_abs_ss[e][iv]=t[e].tau_absorber_final;
	  m->S0[e][iv]=t[e].holstein_T_final;
	}
      }
    }
    free(b); free(infl); free(touched);
  }
  return steps;
}

/* emission_voxels::solve emission_voxels.hpp:170-176 + singlet_CFR::pre_solve :402-404:
   (I - branching*K) S = S0 by LU with partial pivoting (Eigen 3.4.0 PartialPivLU in the
   reference -- third-party, not under /root/reference; this is the textbook algorithm).
   K is left untouched here (the reference scales it in place). Returns the relative
   residual max|AS-S0|/max|S0|. */
double oracle_solve(void *h, int e) {
  omodel *m=(omodel*) h;
  const int N=m->n_vox;
  double *A=(double*) malloc(sizeof(double)*(size_t)N*N);
  double *x=(double*) malloc(sizeof(double)*N);
#ifdef ORACLE_FLOAT
  /* float reference solves in float; we mirror the arithmetic type */
  float *Af=(float*) malloc(sizeof(float)*(size_t)N*N); float *xf=(float*) malloc(sizeof(float)*N);
  for (int i=0;i<N;i++) { for (int j=0;j<N;j++) { float k=m->K[e][(size_t)i*N+j]*m->branching[e]; Af[(size_t)i*N+j]=(i==j ? 1.0f : 0.0f)-k; } xf[i]=m->S0[e][i]; }
  for (int k=0;k<N;k++) {
    int p=k; float best=fabsf(Af[(size_t)k*N+k]);
    for (int i=k+1;i<N;i++) { float vv=fabsf(Af[(size_t)i*N+k]); if (vv>best) {best=vv;p=i;} }
    if (p!=k) { for (int j=0;j<N;j++) { float tt=Af[(size_t)k*N+j]; Af[(size_t)k*N+j]=Af[(size_t)p*N+j]; Af[(size_t)p*N+j]=tt; } float tt=xf[k]; xf[k]=xf[p]; xf[p]=tt; }
    float inv=1.0f/Af[(size_t)k*N+k];
    for (int i=k+1;i<N;i++) { float f=Af[(size_t)i*N+k]*inv; if (f!=0) { for (int j=k+1;j<N;j++) Af[(size_t)i*N+j]-=f*Af[(size_t)k*N+j]; xf[i]-=f*xf[k]; } }
  }
  for (int k=N-1;k>=0;k--) { float s=xf[k]; for (int j=k+1;j<N;j++) s-=Af[(size_t)k*N+j]*xf[j]; xf[k]=s/Af[(size_t)k*N+k]; }
  for (int i=0;i<N;i++) m->S[e][i]=xf[i];
  free(Af); free(xf);
#else
  for (int i=0;i<N;i++) { for (int j=0;j<N;j++) { double k=m->K[e][(size_t)i*N+j]*m->branching[e]; A[(size_t)i*N+j]=(i==j ? 1.0 : 0.0)-k; } x[i]=m->S0[e][i]; }
  for (int k=0;k<N;k++) {
    int p=k; double best=fabs(A[(size_t)k*N+k]);
    for (int i=k+1;i<N;i++) { double vv=fabs(A[(size_t)i*N+k]); if (vv>best) {best=vv;p=i;} }
    if (p!=k) { for (int j=0;j<N;j++) { double tt=A[(size_t)k*N+j]; A[(size_t)k*N+j]=A[(size_t)p*N+j]; A[(size_t)p*N+j]=tt; } double tt=x[k]; x[k]=x[p]; x[p]=tt; }
    double inv=1.0/A[(size_t)k*N+k];
#pragma omp parallel for schedule(static)
    for (int i=k+1;i<N;i++) { double f=A[(size_t)i*N+k]*inv; if (f!=0) { double *ri=A+(size_t)i*N; const double *rk=A+(size_t)k*N; for (int j=k+1;j<N;j++) ri[j]-=f*rk[j]; x[i]-=f*x[k]; } }
  }
  for (int k=N-1;k>=0;k--) { double s=x[k]; for (int j=k+1;j<N;j++) s-=A[(size_t)k*N+j]*x[j]; x[k]=s/A[(size_t)k*N+k]; }
  for (int i=0;i<N;i++) m->S[e][i]=(REAL) x[i];
#endif
  double rmax=0, smax=0;
  for (int i=0;i<N;i++) {
    double acc=0;
    for (int j=0;j<N;j++) acc += ((i==j ? 1.0 : 0.0)-(double) m->K[e][(size_t)i*N+j]*(double) m->branching[e])*(double) m->S[e][j];
    double rr=fabs(acc-(double) m->S0[e][i]); if (rr>rmax) rmax=rr;
    if (fabs((double) m->S0[e][i])>smax) smax=fabs((double) m->S0[e][i]);
  }
  free(A); free(x);
  return rmax/(smax>0 ? smax : 1.0);
}

void oracle_get_K(void *h, int e, double *out) {
  omodel *m=(omodel*) h; size_t n=(size_t) m->n_vox*m->n_vox;
  for (size_t i=0;i<n;i++) out[i]=m->K[e][i];
}
void oracle_get_vectors(void *h, int e, double *S0, double *tsp, double *tab, double *S) {
  omodel *m=(omodel*) h;
  for (int i=0;i<m->n_vox;i++) { S0[i]=m->S0[e][i]; tsp[i]=m->tau_sp_ss[e][i]; tab[i]=m->tau_abs_ss[e][i]; S[i]=m->S[e][i]; }
}
void oracle_set_sourcefn(void *h, int e, const double *S) {
  omodel *m=(omodel*) h;
  for (int i=0;i<m->n_vox;i++) m->S[e][i]=(REAL) S[i];
}

/* atmo_vector::extend atmo_vec.cpp:292-306 (+ atmo_point::xyz, operator*) -> r, t of the point */
static void o_extend(const ovec *v, REAL dist, REAL *r_out, REAL *t_out) {
  const REAL scale = (REAL) 1e9;
  REAL newx = v->x/scale;
  newx += (v->lx*dist)/scale;
  const REAL newy = v->y/scale + (v->ly*dist)/scale;
  const REAL newz = v->z/scale + (v->lz*dist)/scale;
  REAL r = hypot(hypot(newx,newy),newz);
  REAL t = acos(newz/r);
  *r_out = r*scale;
  *t_out = t;
}

/* spherical_azimuthally_symmetric_grid::interp_weights :511-612 */
static void o_interp_weights(const omodel *m, int ivoxel, REAL r, REAL t, int *idx, REAL *w) {
  int r_idx, sza_idx;
  if (ivoxel<0 || ivoxel>m->n_vox-1) { r_idx=-1; sza_idx=-1; }
  else { r_idx=ivoxel/(m->n_sb-1); sza_idx=ivoxel%(m->n_sb-1); }
  const REAL *rb=m->rb, *sb=m->sb;
  if (r < rb[r_idx] && rb[r_idx]/r > (1-O_EPS)) r = rb[r_idx]+O_EPS;
  if (rb[r_idx+1] < r && r/rb[r_idx+1] < (1+O_EPS)) r = rb[r_idx+1]-O_EPS;
  if (t < sb[sza_idx] && sb[sza_idx]/t > (1-O_CONEEPS)) t = sb[sza_idx]+O_CONEEPS;
  if (sb[sza_idx+1] < t && t/sb[sza_idx+1] < (1+O_CONEEPS)) t = sb[sza_idx+1]-O_CONEEPS;
  int rlo, rhi; REAL r_wt;
  if (r_idx==0 && r <= m->pts_r[0]) { rlo=rhi=0; r_wt=1.0; }
  else if (r_idx==m->n_rb-2 && m->pts_r[m->n_rb-2] <= r) { rlo=rhi=m->n_rb-2; r_wt=0.0; }
  else {
    if (r < m->pts_r[r_idx]) { rlo=r_idx-1; rhi=rlo+1; } else { rlo=r_idx; rhi=rlo+1; }
    r_wt = (STD_LOG(r)-m->log_pts_r[rlo])/(m->log_pts_r[rhi]-m->log_pts_r[rlo]);
  }
  int slo, shi; REAL s_wt;
  if (t < m->pts_s[sza_idx]) { slo=sza_idx-1; shi=slo+1; } else { slo=sza_idx; shi=slo+1; }
  s_wt = (t-m->pts_s[slo])/(m->pts_s[shi]-m->pts_s[slo]);
  idx[0]=o_vox(m,rlo,slo); w[0]=(RL(1.0)-r_wt)*(RL(1.0)-s_wt);
  idx[1]=o_vox(m,rhi,slo); w[1]=r_wt*(RL(1.0)-s_wt);
  idx[2]=o_vox(m,rlo,shi); w[2]=(RL(1.0)-r_wt)*s_wt;
  idx[3]=o_vox(m,rhi,shi); w[3]=r_wt*s_wt;
}

static REAL o_interp(const REAL *q, const int *idx, const REAL *w) {   /* emission_voxels.hpp:58-70 */
  REAL s=0;
  for (int k=0;k<4;k++) s += w[k]*q[idx[k]];
  return s;
}

/* RT_grid::brightness(vec, los, n_subsamples) RT_grid.hpp:233-299;
   emission_voxels::update_tracker_brightness_{interp,nointerp} :199-233;
   singlet_CFR::update_tracker_start_interp :315-348, update_tracker_brightness :262-276.
   out[n_em][4][n_los]: brightness, tau_species_final, tau_absorber_final, species_col_dens */
void oracle_brightness(void *h, int n_los, const double *loc, const double *dir, int n_subsamples, double *out) {
  omodel *m=(omodel*) h;
#pragma omp parallel
  {
    obnd *b=(obnd*) malloc(sizeof(obnd)*m->cap);
#pragma omp for schedule(dynamic,64)
    for (int i=0;i<n_los;i++) {
      ovec v; o_los(loc+3*i, dir+3*i, &v);
      int begin, eb; int n=o_traverse(m,&v,b,&begin,&eb);
      otracker t[2];
      for (int e=0;e<m->n_em;e++) o_reset(&t[e], 0.0);
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
	    int idx[4]; REAL w[4];
	    if (n_subsamples!=0) o_interp_weights(m,cur,pr,pt,idx,w);
	    for (int e=0;e<m->n_em;e++) {
	      REAL **a=m->arr[e]; REAL Sv;
	      if (n_subsamples==0) {
		o_update(0, a[A_TR][cur], a[A_N][cur], a[A_DTS][cur], a[A_DTA][cur], d_step, &t[e]);
		Sv=m->S[e][cur];
	      } else {
		REAL Tr=o_interp(a[A_TR_PT],idx,w);
		REAL dn=o_interp(a[A_N_PT],idx,w);
		REAL ds=o_interp(a[A_DTS_PT],idx,w);
		REAL da=o_interp(a[A_DTA_PT],idx,w);
		o_update(0, Tr, dn, ds, da, d_step, &t[e]);
		Sv=o_interp(m->S[e],idx,w);
	      }
	      t[e].brightness += (Sv*m->g_factor[e]*m->branching[e]*t[e].holstein_T_int
				  /m->sigma_ref[e]*O_ONE_OVER_SQRT_PI/RL(1e9));
	    }
	  }
	}
	if (eb) for (int e=0;e<m->n_em;e++) t[e].tau_absorber_final=-1.0;
      }
      for (int e=0;e<m->n_em;e++) {
	out[((size_t)e*4+0)*n_los+i]=t[e].brightness;
	out[((size_t)e*4+1)*n_los+i]=t[e].tau_species_final;
	out[((size_t)e*4+2)*n_los+i]=t[e].tau_absorber_final;
	out[((size_t)e*4+3)*n_los+i]=t[e].species_col_dens;
      }
    }
    free(b);
  }
}

/* ------------------------------------------------------------------ multiplet CFR emissions */
#include "multiplet_oracle.inc.c"
static omstate *om_get(omodel *m) {
  if (!m->mult_state) m->mult_state=calloc(1, sizeof(omstate));
  return (omstate*) m->mult_state;
}
static void om_free(omodel *m) {
  omstate *s=(omstate*) m->mult_state;
  if (!s) return;
  for (int l=0;l<M_MAXLOW;l++) { free(s->n[l]); free(s->n_pt[l]); }
  free(s->T); free(s->T_pt); free(s->nabs); free(s->nabs_pt);
  free(s->K); free(s->S0); free(s->tsp); free(s->tab); free(s->S);
  free(s);
  m->mult_state=NULL;
}
