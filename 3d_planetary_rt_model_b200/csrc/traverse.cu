// traverse.cu -- ray / voxel-boundary traversal on the device (sm_100a).
//
// COMPILE THIS FILE WITH -fmad=false AND WITHOUT --use_fast_math: the integer
// voxel sequence it produces must be bit-identical to the reference CPU build
// (x86-64, no FMA contraction), so every product/sum is rounded separately and
// sqrt / division are the IEEE-rounded ones.  Transcendentals never appear here:
// they are tabulated on the host (grid_host.cpp).
//
// What it computes (reference src/, restated; nothing is shared with RT_gpu.cu):
//   spherical_azimuthally_symmetric_grid::ray_voxel_intersections
//       grid/grid_spherical_azimuthally_symmetric.hpp:459-509
//   plane_parallel_grid::ray_voxel_intersections     grid/grid_plane_parallel.hpp:268-302 (GridView::pp)
//   sphere:: / cone:: / plane::intersections         grid/intersections.cpp:58-95, 120-164, 25-46
//   boundary_set::add_intersections / sort / propagate_indices /
//       assign_voxel_indices / trim                  grid/boundaries.hpp:131-232
//   boundary_intersection_stepper::init_stepper      grid/boundaries.hpp:334-349
//
// Mapping: one warp per ray.  The reference builds the full crossing list
// (<= 2*n_rb+n_sb entries), insertion-sorts it and then trims it to the first
// contiguous in-grid run.  Here the lanes evaluate the primitives in parallel and
// the trim is applied BEFORE the sort: a list entry can only be "outside the grid"
// through its radial index, so the first in-grid entry and the first later exit are
// found from the sphere crossings alone (two warp min-reductions over
// (distance, insertion-slot) keys); only crossings between those two keys are
// compacted into shared memory, ranked (stable order = (distance, slot), which is
// what a stable insertion sort with strict '<' yields) and forward-filled.
// The output is the trimmed boundary list of the reference, entry for entry.
#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include "common.hpp"

namespace b200rt {

namespace {

template <class Real> struct Lim;
template <> struct Lim<double> {
  __device__ static double strict_eps() { return 1e-10; }   // Real.hpp:24
  __device__ static double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};
template <> struct Lim<float> {
  __device__ static float strict_eps() { return 1e-5f; }    // Real.hpp:15
  __device__ static float inf() { return __int_as_float(0x7f800000); }
};

// info word: bits 0-11 insertion slot, bit 12 dimension (0 radial, 1 sza), bits 13.. value+1
constexpr int SLOT_BITS = 12;
__device__ __forceinline__ int pack_info(int slot, int dim, int val) {
  return slot | (dim << SLOT_BITS) | ((val + 1) << (SLOT_BITS + 1));
}
__device__ __forceinline__ int info_slot(int info) { return info & ((1 << SLOT_BITS) - 1); }
__device__ __forceinline__ int info_dim(int info) { return (info >> SLOT_BITS) & 1; }
__device__ __forceinline__ int info_val(int info) { return (info >> (SLOT_BITS + 1)) - 1; }

// 16-byte groups of distances for the ranking loop
template <class Real> struct VecOf;
template <> struct VecOf<double> {
  typedef double2 type;
  __device__ static __forceinline__ int count_less(const double2 &v, double d) { return (v.x < d ? 1 : 0) + (v.y < d ? 1 : 0); }
};
template <> struct VecOf<float> {
  typedef float4 type;
  __device__ static __forceinline__ int count_less(const float4 &v, float d) {
    return (v.x < d ? 1 : 0) + (v.y < d ? 1 : 0) + (v.z < d ? 1 : 0) + (v.w < d ? 1 : 0);
  }
};

template <class Real>
__device__ __forceinline__ bool key_less(Real d1, int s1, Real d2, int s2) {
  return d1 < d2 || (d1 == d2 && s1 < s2);
}

template <class Real>
__device__ __forceinline__ void warp_min_key(Real &d, int &s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Real d2 = __shfl_xor_sync(0xffffffffu, d, o);
    int s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (key_less(d2, s2, d, s)) { d = d2; s = s2; }
  }
}
template <class Real>
__device__ __forceinline__ void warp_max_key(Real &d, int &s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Real d2 = __shfl_xor_sync(0xffffffffu, d, o);
    int s2 = __shfl_xor_sync(0xffffffffu, s, o);
    if (key_less(d, s, d2, s2)) { d = d2; s = s2; }
  }
}

__device__ __forceinline__ bool samesign_f(double a, double b) {
  return (a > 0 && b > 0) || (a < 0 && b < 0) || (a == 0 && b == 0);
}

// sphere::intersections + the ordering of add_intersections: first <= second, +inf = absent
template <class Real>
__device__ __forceinline__ int sphere_hits(Real r, Real cost, Real R2, Real &first, Real &second) {
  const Real scale = Real(1e9);
  const Real r_norm = r / scale;
  const Real B = r_norm * cost;
  const Real C = r_norm * r_norm - R2;
  const Real discr = B * B - C;
  Real dd[2];
  int nh = 0;
  if (discr > 0) {
    const Real sq = sqrt(discr);
    const Real d0 = (B > 0) ? -B - sq : -B + sq;
    if (d0 > 0) { dd[nh] = d0 * scale; nh++; }
    const Real d1 = C / d0;
    if (d1 > 0) { dd[nh] = d1 * scale; nh++; }
  }
  first = second = Lim<Real>::inf();
  if (nh == 1) first = dd[0];
  else if (nh == 2) {
    const bool in_order = dd[1] > dd[0];
    first = in_order ? dd[0] : dd[1];
    second = in_order ? dd[1] : dd[0];
  }
  return nh;
}

// plane::intersections (grid/intersections.cpp:25-46): the plane z = zb, at most one hit
template <class Real>
__device__ __forceinline__ int plane_hits(Real z, Real lz, Real zb, Real &first, Real &second) {
  first = second = Lim<Real>::inf();
  if (lz != 0) {
    const Real d = (zb - z) / lz;
    if (d > 0) { first = d; return 1; }
  }
  return 0;
}

template <class Real>
__device__ __forceinline__ int cone_hits(Real r, Real zn, Real lz, Real cost, Real ca, Real ca2, Real &first,
                                         Real &second) {
  const Real A = lz * lz - ca2;
  const Real B = zn * lz - cost * ca2;
  const Real C = zn * zn - ca2;
  Real dd[2];
  int nh = 0;
  const Real tol = Lim<Real>::strict_eps();
  if (A > tol || A < -tol) {
    const Real discr = B * B - A * C;
    if (discr > 0) {
      const Real sq = sqrt(discr);
      const Real q = (B > 0) ? -B - sq : -B + sq;
      const Real d0 = q / A;
      if (d0 > 0 && samesign_f(zn + d0 * lz, ca)) { dd[nh] = d0 * r; nh++; }
      const Real d1 = C / q;
      if (d1 > 0 && samesign_f(zn + d1 * lz, ca)) { dd[nh] = d1 * r; nh++; }
    }
  } else {
    const Real d = -C / (2 * B);
    if (d > 0 && samesign_f(zn + d * lz, ca)) { dd[nh] = d * r; nh++; }
  }
  first = second = Lim<Real>::inf();
  if (nh == 1) first = dd[0];
  else if (nh == 2) {
    const bool in_order = dd[1] > dd[0];
    first = in_order ? dd[0] : dd[1];
    second = in_order ? dd[1] : dd[0];
  }
  return nh;
}

// find_coordinate_index (grid_spherical_azimuthally_symmetric.hpp:433-450): index of the
// first boundary with c < b[i], minus one (n-1 if none); warp-cooperative.
template <class Real>
__device__ __forceinline__ int find_index(Real c, const Real *__restrict__ b, int n, int lane) {
  for (int base = 0; base < n; base += 128) {
    // four groups of 32 boundaries with their loads in flight together
    unsigned m[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int i = base + 32 * q + lane;
      const bool p = (i < n) && (c < b[i]);
      m[q] = __ballot_sync(0xffffffffu, p);
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
      if (m[q]) return base + 32 * q + __ffs(m[q]) - 2;
  }
  return n - 1;
}

// ---- the general path: any origin (inside or outside the grid), any primitive set (spheres + cones, or planes).
// One warp, one ray; per-warp scratch `base` (general_scratch_bytes).  The trim is applied before the ordering, the
// survivors are ranked by (distance, slot).
template <class Real>
__device__ __forceinline__ size_t general_scratch_bytes(int n_rb, int cap) {
  const int cap4 = (cap + 3) & ~3;
  return ((size_t) (2 * n_rb + 2 * cap4) * sizeof(Real) + (size_t) 2 * cap * sizeof(int) + 15) & ~size_t(15);
}

template <class Real>
__device__ __noinline__ void general_ray(const GridView<Real> &g, unsigned char *base, long long ray, Real r, Real z, Real t,
                                         Real cost, Real lz, int r0, int s0, bool origin_in, const ListView<Real> &out,
                                         int *overflow_flag) {
  const int lane = threadIdx.x & 31;
  const int cap = g.cap, n_rb = g.n_rb, n_sb = g.n_sb;
  const int cap4 = (cap + 3) & ~3;                         // room for the 16-byte padding of the ranking loop
  Real *cmp_d = reinterpret_cast<Real *>(base);            // [cap4]    compacted, unsorted (16-byte aligned)
  Real *srt_d = cmp_d + cap4;                              // [cap4]    sorted
  Real *sph_d = srt_d + cap4;                              // [2*n_rb]  first/second hit per sphere
  int *cmp_i = reinterpret_cast<int *>(sph_d + 2 * n_rb);  // [cap]
  int *srt_i = cmp_i + cap;                                // [cap]
  const Real INF = Lim<Real>::inf();

  // ---- pass 1: spheres -> shared, and the keys of the first in-grid entry / first exit
#pragma unroll 1
  for (int ir = lane; ir < n_rb; ir += 32) {
    Real f, s;
    if (g.pp) plane_hits(z, lz, g.rb[ir], f, s);   // plane_parallel_grid::ray_voxel_intersections (grid_plane_parallel.hpp:282-288)
    else sphere_hits(r, cost, g.sph_R2[ir], f, s);
    sph_d[2 * ir] = f;
    sph_d[2 * ir + 1] = s;
  }
  __syncwarp();

  Real db = INF; int sb_slot = 0x7fffffff;      // key of `begin`
  int rb_val = r0;                               // radial index at `begin`
  if (origin_in) { db = 0; sb_slot = 0; }
  else {
#pragma unroll 1
    for (int ir = lane; ir < n_rb; ir += 32) {
      const bool above = r > g.rb[ir];
      const Real f = sph_d[2 * ir], s = sph_d[2 * ir + 1];
      const int vf = above ? ir - 1 : ir;        // value set by the first hit
      const int vs = above ? ir : ir - 1;        // value set by the second hit (2 hits only)
      if (f < INF && vf >= 0 && vf <= n_rb - 2 && key_less(f, 1 + 2 * ir, db, sb_slot)) { db = f; sb_slot = 1 + 2 * ir; }
      if (s < INF && vs >= 0 && vs <= n_rb - 2 && key_less(s, 2 + 2 * ir, db, sb_slot)) { db = s; sb_slot = 2 + 2 * ir; }
    }
    warp_min_key(db, sb_slot);
    if (db < INF) {
      const int ir = (sb_slot - 1) >> 1;
      const bool above = r > g.rb[ir];
      const bool is_first = ((sb_slot - 1) & 1) == 0;
      rb_val = is_first ? (above ? ir - 1 : ir) : (above ? ir : ir - 1);
    }
  }
  if (!(db < INF)) {   // never inside the grid
    if (lane == 0) { out.len[ray] = 0; out.flag[ray] = 0; }
    return;
  }
  Real de = INF; int se_slot = 0x7fffffff;       // key of `end`
#pragma unroll 1
  for (int ir = lane; ir < n_rb; ir += 32) {
    const bool above = r > g.rb[ir];
    const Real f = sph_d[2 * ir], s = sph_d[2 * ir + 1];
    const int vf = above ? ir - 1 : ir;
    const int vs = above ? ir : ir - 1;
    if (f < INF && (vf < 0 || vf > n_rb - 2) && key_less(db, sb_slot, f, 1 + 2 * ir) && key_less(f, 1 + 2 * ir, de, se_slot)) { de = f; se_slot = 1 + 2 * ir; }
    if (s < INF && (vs < 0 || vs > n_rb - 2) && key_less(db, sb_slot, s, 2 + 2 * ir) && key_less(s, 2 + 2 * ir, de, se_slot)) { de = s; se_slot = 2 + 2 * ir; }
  }
  warp_min_key(de, se_slot);

  // ---- pass 2: compact every crossing with  begin < key <= end
  int count = 0;                                  // warp-uniform
  Real dsb = -INF; int ssb_slot = -1; int sb_val = s0;   // latest sza crossing before `begin`
#pragma unroll 1
  for (int base_ir = 0; base_ir < n_rb; base_ir += 32) {
    const int ir = base_ir + lane;
    Real f = INF, s = INF;
    bool above = false;
    if (ir < n_rb) { f = sph_d[2 * ir]; s = sph_d[2 * ir + 1]; above = r > g.rb[ir]; }
    const bool kf = f < INF && key_less(db, sb_slot, f, 1 + 2 * ir) && !key_less(de, se_slot, f, 1 + 2 * ir);
    const bool ks = s < INF && key_less(db, sb_slot, s, 2 + 2 * ir) && !key_less(de, se_slot, s, 2 + 2 * ir);
    const unsigned mf = __ballot_sync(0xffffffffu, kf);
    const unsigned ms = __ballot_sync(0xffffffffu, ks);
    const unsigned lt = (1u << lane) - 1u;
    if (kf) {
      const int pos = count + __popc(mf & lt);
      if (pos < cap - 1) { cmp_d[pos] = f; cmp_i[pos] = pack_info(1 + 2 * ir, 0, above ? ir - 1 : ir); }
    }
    count += __popc(mf);
    if (ks) {
      const int pos = count + __popc(ms & lt);
      if (pos < cap - 1) { cmp_d[pos] = s; cmp_i[pos] = pack_info(2 + 2 * ir, 0, above ? ir : ir - 1); }
    }
    count += __popc(ms);
  }
  const Real zn = z / r;
#pragma unroll 1
  for (int base_k = 0; base_k < n_sb - 2; base_k += 32) {
    const int k = base_k + lane;
    Real f = INF, s = INF;
    bool above = false;
    if (k < n_sb - 2) {
      cone_hits(r, zn, lz, cost, g.cone_cos[k], g.cone_cos2[k], f, s);
      above = t > g.sb[k + 1];
    }
    const int slot_f = 1 + 2 * n_rb + 2 * k, slot_s = slot_f + 1;
    const int vf = above ? k : k + 1;            // idx = k+1: above ? idx-1 : idx
    const int vs = above ? k + 1 : k;
    const bool kf = f < INF && key_less(db, sb_slot, f, slot_f) && !key_less(de, se_slot, f, slot_f);
    const bool ks = s < INF && key_less(db, sb_slot, s, slot_s) && !key_less(de, se_slot, s, slot_s);
    if (!origin_in) {   // remember the latest sza crossing strictly before `begin`
      if (f < INF && key_less(f, slot_f, db, sb_slot) && key_less(dsb, ssb_slot, f, slot_f)) { dsb = f; ssb_slot = slot_f; sb_val = vf; }
      if (s < INF && key_less(s, slot_s, db, sb_slot) && key_less(dsb, ssb_slot, s, slot_s)) { dsb = s; ssb_slot = slot_s; sb_val = vs; }
    }
    const unsigned mf = __ballot_sync(0xffffffffu, kf);
    const unsigned ms = __ballot_sync(0xffffffffu, ks);
    const unsigned lt = (1u << lane) - 1u;
    if (kf) {
      const int pos = count + __popc(mf & lt);
      if (pos < cap - 1) { cmp_d[pos] = f; cmp_i[pos] = pack_info(slot_f, 1, vf); }
    }
    count += __popc(mf);
    if (ks) {
      const int pos = count + __popc(ms & lt);
      if (pos < cap - 1) { cmp_d[pos] = s; cmp_i[pos] = pack_info(slot_s, 1, vs); }
    }
    count += __popc(ms);
  }
  int s_begin = s0;
  if (!origin_in) {
    // value carried by the max-key sza crossing before `begin` (warp reduce)
    Real dmax = dsb; int smax = ssb_slot;
    warp_max_key(dmax, smax);
    // the lane owning that key broadcasts its value
    const unsigned own = __ballot_sync(0xffffffffu, smax >= 0 && ssb_slot == smax && dsb == dmax);
    if (own) s_begin = __shfl_sync(0xffffffffu, sb_val, __ffs(own) - 1);
  }
  if (count + 1 > cap) {   // hard capacity check (the reference only asserts, boundaries.hpp:153-158)
    if (lane == 0) { out.len[ray] = 0; out.flag[ray] = 2; atomicExch(overflow_flag, 1); }
    return;
  }
  if (count == 0) {        // `begin` is the last entry of the list: empty (boundaries.hpp:219-220)
    if (lane == 0) { out.len[ray] = 0; out.flag[ray] = 0; }
    return;
  }
  __syncwarp();

  // ---- rank by (distance, slot): what the stable insertion sort produces.
  // Fast path: rank by distance alone (one compare per pair, distances read 16 bytes at a time); two crossings at
  // exactly the same distance then collide on one rank and leave a slot of the sorted list unwritten, which is
  // detected below and sends the (rare) ray through the exact (distance, slot) ranking.
  constexpr int VEC = 16 / (int) sizeof(Real);
  typedef typename VecOf<Real>::type RealV;
  const int count_pad = (count + VEC - 1) / VEC * VEC;
#pragma unroll 1
  for (int e = count + lane; e < count_pad; e += 32) cmp_d[e] = INF;     // padding never counts (INF < d is false)
#pragma unroll 1
  for (int e = lane; e < count; e += 32) srt_i[e] = -1;
  __syncwarp();
#pragma unroll 1
  for (int e = lane; e < count; e += 32) {
    const Real d = cmp_d[e];
    int rank = 0;
    const RealV *cv = reinterpret_cast<const RealV *>(cmp_d);
#pragma unroll 4
    for (int j = 0; j < count_pad / VEC; j++) rank += VecOf<Real>::count_less(cv[j], d);
    srt_d[rank] = d;
    srt_i[rank] = cmp_i[e];
  }
  __syncwarp();
  bool hole = false;
#pragma unroll 1
  for (int e = lane; e < count; e += 32) hole |= (srt_i[e] < 0);
  if (__any_sync(0xffffffffu, hole)) {
    __syncwarp();
#pragma unroll 1
    for (int e = lane; e < count; e += 32) {
      const Real d = cmp_d[e];
      const int info = cmp_i[e];
      const int slot = info_slot(info);
      int rank = 0;
      for (int j = 0; j < count; j++) {
        const Real dj = cmp_d[j];
        const int sj = info_slot(cmp_i[j]);
        rank += key_less(dj, sj, d, slot) ? 1 : 0;
      }
      srt_d[rank] = d;
      srt_i[rank] = info;
    }
  }
  __syncwarp();

  // ---- forward fill (propagate_indices), voxel ids (assign_voxel_indices), write out
  const int len = count + 1;
  Real *od = out.dist + (size_t) ray * cap;
  int *oe = out.ent + (size_t) ray * cap;
  int carry_r = rb_val, carry_s = s_begin;
  if (lane == 0) {
    od[0] = db;
    oe[0] = carry_r * (n_sb - 1) + carry_s;      // `begin` is inside the grid by construction
  }
  int last_r = carry_r;
#pragma unroll 1
  for (int base_e = 0; base_e < count; base_e += 32) {
    const int e = base_e + lane;
    int pr = 0, ps = 0;                           // (pos+1)<<16 | (val+1) ; 0 = not set here
    Real d = 0;
    if (e < count) {
      const int info = srt_i[e];
      d = srt_d[e];
      const int packed = ((lane + 1) << 16) | (info_val(info) + 1);
      if (info_dim(info) == 0) pr = packed; else ps = packed;
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int qr = __shfl_up_sync(0xffffffffu, pr, o);
      const int qs = __shfl_up_sync(0xffffffffu, ps, o);
      if (lane >= o) { pr = max(pr, qr); ps = max(ps, qs); }
    }
    const int rv = pr ? (pr & 0xffff) - 1 : carry_r;
    const int sv = ps ? (ps & 0xffff) - 1 : carry_s;
    if (e < count) {
      const bool inside = rv >= 0 && rv <= n_rb - 2 && sv >= 0 && sv <= n_sb - 2;
      od[e + 1] = d;
      oe[e + 1] = inside ? rv * (n_sb - 1) + sv : -1;
    }
    const int last_lane = min(31, count - base_e - 1);
    carry_r = __shfl_sync(0xffffffffu, rv, last_lane);
    carry_s = __shfl_sync(0xffffffffu, sv, last_lane);
    last_r = carry_r;
  }
  if (lane == 0) {
    out.len[ray] = len;
    out.flag[ray] = (last_r == -1) ? 1 : 0;       // exits_bottom (boundaries.hpp:340)
  }
  __syncwarp();
}

template <class Real, bool VOXEL_RAYS>
__device__ __forceinline__ void ray_scalars(const GridView<Real> &g, const RayList<Real> &rl, int v_begin, long long ray,
                                            int lane, Real &r, Real &z, Real &t, Real &cost, Real &lz, int &r0, int &s0) {
  const int n_rb = g.n_rb, n_sb = g.n_sb;
  int i_voxel;
  if (VOXEL_RAYS) {
    // a batch of voxel-origin rays is far below 2^32 rays (launch_impl checks): 32-bit division
    const unsigned slot_v = (unsigned) ray / (unsigned) g.n_rays;
    int iv = v_begin + (int) slot_v;
    if (g.vox_map) iv = g.vox_map[iv];
    const int ir = (int) ((unsigned) ray - slot_v * (unsigned) g.n_rays);
    const int irad = iv / (n_sb - 1), isza = iv % (n_sb - 1);
    r = g.pts_r[irad];
    t = g.pts_s[isza];
    z = g.vox_z[iv];
    cost = g.ray_cost[ir];
    // line_z = ray.cost*cos(pt.t) - cos(ray.p)*ray.sint*sin(pt.t)   (atmo_vec.cpp:246), in double
    const double a = (double) cost * g.col_ct[isza];
    const double b = (g.ray_cp[ir] * (double) g.ray_sint[ir]) * g.col_st[isza];
    lz = (Real) (a - b);
    i_voxel = iv;
  } else {
    r = rl.r[ray]; z = rl.z[ray]; t = rl.t[ray]; cost = rl.cost[ray]; lz = rl.lz[ray];
    i_voxel = rl.i_voxel ? rl.i_voxel[ray] : -1;
  }
  // ---- origin entry (:467-477)
  if (i_voxel == -1) {
    r0 = find_index(r, g.rb, n_rb, lane);
    s0 = find_index(t, g.sb, n_sb, lane);
  } else {
    r0 = i_voxel / (n_sb - 1);
    s0 = i_voxel % (n_sb - 1);
  }
}

// every ray through the general path: plane-parallel grids, grids with more than 128 radial boundaries
template <class Real, bool VOXEL_RAYS>
__global__ void __launch_bounds__(128)
traverse_kernel(GridView<Real> g, int v_begin, long long n_total, RayList<Real> rl, ListView<Real> out,
                int *overflow_flag) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  unsigned char *base = smem_raw + general_scratch_bytes<Real>(g.n_rb, g.cap) * warp;
  for (long long ray = (long long) blockIdx.x * warps_per_block + warp; ray < n_total;
       ray += (long long) gridDim.x * warps_per_block) {
    Real r, z, t, cost, lz;
    int r0, s0;
    ray_scalars<Real, VOXEL_RAYS>(g, rl, v_begin, ray, lane, r, z, t, cost, lz, r0, s0);
    const bool origin_in = (r0 >= 0 && r0 <= g.n_rb - 2 && s0 >= 0 && s0 <= g.n_sb - 2);
    general_ray<Real>(g, base, ray, r, z, t, cost, lz, r0, s0, origin_in, out, overflow_flag);
  }
}

// ---- the fast path: rays that START INSIDE the spherical grid (every voxel-origin ray; lines of sight of a spacecraft
// inside the model domain).  What it exploits, and how the result stays the reference's list entry for entry:
//   * along a ray the sphere crossings come in three monotone runs -- inner spheres entered in descending radius, left
//     again in ascending radius, then the outer spheres in ascending radius -- so every sphere crossing knows its place
//     in the ordered sphere list from three counts (ballots), without a single comparison.  The list built that way is
//     then VERIFIED (no holes, strictly increasing (distance, slot) keys); a ray that fails -- equal distances out of
//     slot order, a sphere grazed in floating point -- takes the general path, which ranks exactly;
//   * the first exit from the grid (`end` of boundary_set::trim) is a sphere crossing: the first entry of the ordered
//     sphere list whose radial index leaves [0, n_rb-2];
//   * the cone crossings before `end` are few: they are ranked among themselves by (distance, slot) exactly;
//   * the two ordered lists are merged by binary search (a sphere precedes a cone at equal distance: its slot is lower),
//     and since the position of an entry in the OTHER list is then known, so is the index it inherits from it
//     (propagate_indices): every entry writes its final list position, distance and voxel id directly -- no ranking of
//     the whole list (count^2 / 32 compares), no forward-fill scan, no sorted copy in shared memory.
// element-sized asynchronous copy global -> shared (LDGSTS): issued before the cone loop, waited for after it
template <class T>
__device__ __forceinline__ void cp_async_elem(T *smem, const T *gmem) {
  const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
}

template <class Real>
__device__ __forceinline__ int count_below(const Real *a, int n, Real d) {      // #{a[i] < d}, a ascending
  int lo = 0;
  while (n > 0) {
    const int half = n >> 1;
    if (a[lo + half] < d) { lo += half + 1; n -= half + 1; }
    else n = half;
  }
  return lo;
}
template <class Real>
__device__ __forceinline__ int count_not_above(const Real *a, int n, Real d) {  // #{a[i] <= d}, a ascending
  int lo = 0;
  while (n > 0) {
    const int half = n >> 1;
    if (a[lo + half] <= d) { lo += half + 1; n -= half + 1; }
    else n = half;
  }
  return lo;
}

// ordered sphere crossings of a ray that starts inside the grid, trimmed at the first exit (see the fast path below):
// S_d / S_i [0, nS), de = distance of `end` (+inf if the ray never leaves through a sphere).  false: the analytic order
// did not verify -- the caller takes the general path.
template <class Real, int SPH_ITERS>
__device__ __forceinline__ bool build_sphere_list(const GridView<Real> &g, Real r, Real cost, int lane, Real *S_d, int *S_i,
                                                  int &nS, Real &de) {
  const int n_rb = g.n_rb;
  const Real INF = Lim<Real>::inf();
  nS = 0;
  de = INF;
  // ---- spheres: hits in registers, the three counts
  Real hf[SPH_ITERS], hs[SPH_ITERS];
  unsigned above_bits = 0;
  int i0 = 0, n2 = 0, total = 0;
#pragma unroll
  for (int it = 0; it < SPH_ITERS; it++) {
    const int ir = it * 32 + lane;
    hf[it] = INF; hs[it] = INF;
    bool above = false;
    if (ir < n_rb) {
      sphere_hits(r, cost, g.sph_R2[ir], hf[it], hs[it]);
      above = r > g.rb[ir];
    }
    if (above) above_bits |= 1u << it;
    const unsigned mf = __ballot_sync(0xffffffffu, hf[it] < INF), ms = __ballot_sync(0xffffffffu, hs[it] < INF);
    i0 += __popc(__ballot_sync(0xffffffffu, above));
    n2 += __popc(ms);
    total += __popc(mf) + __popc(ms);
  }
  // the place formulas below are a bijection onto [0, total) exactly when the inner spheres that are hit are the n2
  // outermost ones, each hit twice, and every outer sphere is hit once: checked lane by lane
  bool shape_ok = true;
#pragma unroll
  for (int it = 0; it < SPH_ITERS; it++) {
    const int ir = it * 32 + lane;
    if (ir < n_rb) {
      const bool above = (above_bits >> it) & 1u, h1 = hf[it] < INF, h2 = hs[it] < INF;
      shape_ok = shape_ok && (above ? (h2 == (ir >= i0 - n2) && h1 == h2) : (h1 && !h2));
    }
  }
  bool ok = __all_sync(0xffffffffu, shape_ok) && (total == 2 * n2 + (n_rb - i0));
  if (ok) {
#pragma unroll
    for (int it = 0; it < SPH_ITERS; it++) {
      const int ir = it * 32 + lane;
      const bool above = (above_bits >> it) & 1u;
      if (hf[it] < INF) {
        const int pos = above ? (i0 - 1 - ir) : (2 * n2 + ir - i0);
        if (pos >= 0 && pos < total) { S_d[pos] = hf[it]; S_i[pos] = pack_info(1 + 2 * ir, 0, above ? ir - 1 : ir); }
      }
      if (hs[it] < INF) {
        const int pos = above ? (2 * n2 - i0 + ir) : -1;
        if (pos >= 0 && pos < total) { S_d[pos] = hs[it]; S_i[pos] = pack_info(2 + 2 * ir, 0, above ? ir : ir - 1); }
      }
    }
    __syncwarp();
    bool bad = false;
    int first_out = total;
#pragma unroll 1
    for (int j = lane; j < total; j += 32) {
      const int info = S_i[j];
      if (j > 0 && !key_less(S_d[j - 1], info_slot(S_i[j - 1]), S_d[j], info_slot(info))) bad = true;
      const int val = info_val(info);
      if ((val < 0 || val > n_rb - 2) && j < first_out) first_out = j;
    }
    ok = !__any_sync(0xffffffffu, bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_out = min(first_out, __shfl_xor_sync(0xffffffffu, first_out, o));
    if (first_out < total) { nS = first_out + 1; de = S_d[first_out]; }
    else nS = total;
  }
  return ok;
}

// the table of those lists for voxel-origin rays: one warp per (radial shell, polar-angle class)
template <class Real, int SPH_ITERS>
__global__ void __launch_bounds__(128)
sphere_table_kernel(GridView<Real> g, int *hdr, Real *tde, Real *td, int *ti) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nS_max = 2 * g.n_rb;
  Real *S_d = reinterpret_cast<Real *>(smem_raw) + (size_t) warp * nS_max;
  int *S_i = reinterpret_cast<int *>(reinterpret_cast<Real *>(smem_raw) + (size_t) (blockDim.x >> 5) * nS_max) + (size_t) warp * nS_max;
  const int pair = blockIdx.x * (blockDim.x >> 5) + warp;
  if (pair >= (g.n_rb - 1) * g.n_cls) return;
  const Real r = g.pts_r[pair / g.n_cls], cost = g.ray_cost[g.cls_ray[pair % g.n_cls]];
  int nS;
  Real de;
  const bool ok = build_sphere_list<Real, SPH_ITERS>(g, r, cost, lane, S_d, S_i, nS, de);
  __syncwarp();
  if (lane == 0) { hdr[2 * pair] = ok ? nS : -1; hdr[2 * pair + 1] = 0; tde[pair] = de; }
  if (ok)
    for (int j = lane; j < nS; j += 32) { td[(size_t) pair * nS_max + j] = S_d[j]; ti[(size_t) pair * nS_max + j] = S_i[j]; }
}

#ifndef FAST_MIN_BLOCKS
#define FAST_MIN_BLOCKS 6
#endif
template <class Real>
__host__ __device__ __forceinline__ size_t fast_scratch_bytes(int n_rb, int n_sb) {
  return ((size_t) (2 * n_rb + 4 * n_sb) * (sizeof(Real) + sizeof(int)) + 15) & ~size_t(15);
}

template <class Real, bool VOXEL_RAYS, int SPH_ITERS, bool USE_TABLE>
__global__ void __launch_bounds__(128, FAST_MIN_BLOCKS)
traverse_fast_kernel(GridView<Real> g, int v_begin, long long n_total, RayList<Real> rl, ListView<Real> out,
                     int *overflow_flag, unsigned per_warp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int cap = g.cap, n_rb = g.n_rb, n_sb = g.n_sb;
  unsigned char *base = smem_raw + (size_t) per_warp * warp;
  const int nS_max = 2 * n_rb, nC_max = 2 * n_sb;
  Real *S_d = reinterpret_cast<Real *>(base);          // [2 n_rb] ordered sphere crossings
  Real *C_d = S_d + nS_max;                            // [2 n_sb] cone crossings before `end`, as found
  Real *Cs_d = C_d + nC_max;                           // [2 n_sb] the same, ordered
  int *S_i = reinterpret_cast<int *>(Cs_d + nC_max);
  int *C_i = S_i + nS_max;
  int *Cs_i = C_i + nC_max;
  const Real INF = Lim<Real>::inf();
  const unsigned lt = (1u << lane) - 1u;

  for (long long ray = (long long) blockIdx.x * warps_per_block + warp; ray < n_total;
       ray += (long long) gridDim.x * warps_per_block) {
    Real r, z, t, cost, lz;
    int r0, s0;
    ray_scalars<Real, VOXEL_RAYS>(g, rl, v_begin, ray, lane, r, z, t, cost, lz, r0, s0);
    const bool origin_in = (r0 >= 0 && r0 <= n_rb - 2 && s0 >= 0 && s0 <= n_sb - 2);
    Real zn_tab = 0;
    if (VOXEL_RAYS && g.vox_zn) zn_tab = g.vox_zn[r0 * (n_sb - 1) + s0];   // z / r of the voxel point, divided on the host
    if (!origin_in) {   // warp-uniform
      general_ray<Real>(g, base, ray, r, z, t, cost, lz, r0, s0, false, out, overflow_flag);
      continue;
    }

    // ---- ordered, trimmed sphere crossings: shared table (voxel-origin rays) or built here
    bool ok;
    int nS = 0;
    Real de = INF;
    if (USE_TABLE) {
      const unsigned slot_v = (unsigned) ray / (unsigned) g.n_rays;
      const int iv_ = g.vox_map ? g.vox_map[v_begin + (int) slot_v] : v_begin + (int) slot_v;
      const size_t pair = (size_t) (iv_ / (n_sb - 1)) * g.n_cls + g.ray_cls[(int) ((unsigned) ray - slot_v * (unsigned) g.n_rays)];
      nS = g.sph_hdr[2 * pair];
      ok = nS >= 0;
      if (ok) {
        de = g.sph_de[pair];
        const Real *td = g.sph_d + pair * nS_max;
        const int *ti = g.sph_i + pair * nS_max;
#pragma unroll 1
        for (int j = lane; j < nS; j += 32) { cp_async_elem(S_d + j, td + j); cp_async_elem(S_i + j, ti + j); }
      }
    } else {
      ok = build_sphere_list<Real, SPH_ITERS>(g, r, cost, lane, S_d, S_i, nS, de);
    }
    if (!ok) {          // warp-uniform: the exact ranking decides
      __syncwarp();
      general_ray<Real>(g, base, ray, r, z, t, cost, lz, r0, s0, true, out, overflow_flag);
      continue;
    }

    // ---- cones before `end` (a cone crossing AT the distance of `end` comes after it: its slot is higher)
    int nC = 0;
    const Real zn = (VOXEL_RAYS && g.vox_zn) ? zn_tab : z / r;
#pragma unroll 1
    for (int base_k = 0; base_k < n_sb - 2; base_k += 32) {
      const int k = base_k + lane;
      Real f = INF, s = INF;
      bool above = false;
      if (k < n_sb - 2) {
        cone_hits(r, zn, lz, cost, g.cone_cos[k], g.cone_cos2[k], f, s);
        above = t > g.sb[k + 1];
      }
      const int slot_f = 1 + 2 * n_rb + 2 * k, slot_s = slot_f + 1;
      const bool kf = f < de, ks = s < de;
      const unsigned mf = __ballot_sync(0xffffffffu, kf);
      const unsigned ms = __ballot_sync(0xffffffffu, ks);
      if (kf) {
        const int pos = nC + __popc(mf & lt);
        if (pos < nC_max) { C_d[pos] = f; C_i[pos] = pack_info(slot_f, 1, above ? k : k + 1); }
      }
      nC += __popc(mf);
      if (ks) {
        const int pos = nC + __popc(ms & lt);
        if (pos < nC_max) { C_d[pos] = s; C_i[pos] = pack_info(slot_s, 1, above ? k + 1 : k); }
      }
      nC += __popc(ms);
    }
    const int count = nS + nC;
    if (count + 1 > cap || nC > nC_max) {   // hard capacity check (the reference only asserts, boundaries.hpp:153-158)
      if (lane == 0) { out.len[ray] = 0; out.flag[ray] = 2; atomicExch(overflow_flag, 1); }
      __syncwarp();
      continue;
    }
    if (count == 0) {        // `begin` is the last entry of the list: empty (boundaries.hpp:219-220)
      if (lane == 0) { out.len[ray] = 0; out.flag[ray] = 0; }
      __syncwarp();
      continue;
    }
    if (USE_TABLE) asm volatile("cp.async.wait_all;" ::: "memory");     // the shared sphere list has landed
    __syncwarp();
    // cones among themselves: rank by distance alone (distances read 16 bytes at a time); two crossings at exactly the
    // same distance collide on one rank and leave a hole, which sends the (rare) ray through the exact (distance, slot) rank
    {
      constexpr int VEC = 16 / (int) sizeof(Real);
      typedef typename VecOf<Real>::type RealV;
      const int nC_pad = (nC + VEC - 1) / VEC * VEC;
#pragma unroll 1
      for (int e = nC + lane; e < nC_pad; e += 32) C_d[e] = INF;      // padding never counts (INF < d is false)
#pragma unroll 1
      for (int e = lane; e < nC; e += 32) Cs_i[e] = -1;
      __syncwarp();
      const RealV *cv = reinterpret_cast<const RealV *>(C_d);
#pragma unroll 1
      for (int e = lane; e < nC; e += 32) {
        const Real d = C_d[e];
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < nC_pad / VEC; j++) rank += VecOf<Real>::count_less(cv[j], d);
        Cs_d[rank] = d;
        Cs_i[rank] = C_i[e];
      }
      __syncwarp();
      bool hole = false;
#pragma unroll 1
      for (int e = lane; e < nC; e += 32) hole |= (Cs_i[e] < 0);
      if (__any_sync(0xffffffffu, hole)) {
#pragma unroll 1
        for (int e = lane; e < nC; e += 32) {
          const Real d = C_d[e];
          const int info = C_i[e], slot = info_slot(info);
          int rank = 0;
          for (int j = 0; j < nC; j++) rank += key_less(C_d[j], info_slot(C_i[j]), d, slot) ? 1 : 0;
          Cs_d[rank] = d;
          Cs_i[rank] = info;
        }
      }
      __syncwarp();
    }

    // ---- merge: every entry writes its final position, distance and voxel id
    Real *od = out.dist + (size_t) ray * cap;
    int *oe = out.ent + (size_t) ray * cap;
    if (lane == 0) {
      od[0] = Real(0);
      oe[0] = r0 * (n_sb - 1) + s0;
    }
#pragma unroll 1
    for (int j = lane; j < nS; j += 32) {
      const Real d = S_d[j];
      const int nc = count_below(Cs_d, nC, d);
      const int rv = info_val(S_i[j]);
      const int sv = nc ? info_val(Cs_i[nc - 1]) : s0;
      const bool inside = rv >= 0 && rv <= n_rb - 2 && sv >= 0 && sv <= n_sb - 2;
      od[j + nc + 1] = d;
      oe[j + nc + 1] = inside ? rv * (n_sb - 1) + sv : -1;
    }
#pragma unroll 1
    for (int k = lane; k < nC; k += 32) {
      const Real d = Cs_d[k];
      const int ns = count_not_above(S_d, nS, d);
      const int rv = ns ? info_val(S_i[ns - 1]) : r0;
      const int sv = info_val(Cs_i[k]);
      const bool inside = rv >= 0 && rv <= n_rb - 2 && sv >= 0 && sv <= n_sb - 2;
      od[k + ns + 1] = d;
      oe[k + ns + 1] = inside ? rv * (n_sb - 1) + sv : -1;
    }
    if (lane == 0) {
      const int last_r = nS ? info_val(S_i[nS - 1]) : r0;
      out.len[ray] = count + 1;
      out.flag[ray] = (last_r == -1) ? 1 : 0;       // exits_bottom (boundaries.hpp:340)
    }
    __syncwarp();
  }
}

template <class Real>
size_t general_bytes_host(const GridView<Real> &g) {
  const int cap4 = (g.cap + 3) & ~3;
  return ((size_t) (2 * g.n_rb + 2 * cap4) * sizeof(Real) + (size_t) 2 * g.cap * sizeof(int) + 15) & ~size_t(15);
}

template <class Real, bool VR, int SPH_ITERS>
cudaError_t launch_fast(const GridView<Real> &g, int v_begin, long long n_total, RayList<Real> rl, ListView<Real> out,
                        int *overflow_flag, unsigned blocks, int threads, cudaStream_t s) {
  const size_t per_warp = std::max(general_bytes_host(g), fast_scratch_bytes<Real>(g.n_rb, g.n_sb));
  const size_t smem = per_warp * (threads / 32);
  cudaError_t e;
  if (VR && g.sph_hdr) {
    e = cudaFuncSetAttribute(traverse_fast_kernel<Real, VR, SPH_ITERS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    traverse_fast_kernel<Real, VR, SPH_ITERS, true><<<blocks, threads, smem, s>>>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) per_warp);
  } else {
    e = cudaFuncSetAttribute(traverse_fast_kernel<Real, VR, SPH_ITERS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    if (e != cudaSuccess) return e;
    traverse_fast_kernel<Real, VR, SPH_ITERS, false><<<blocks, threads, smem, s>>>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) per_warp);
  }
  return cudaGetLastError();
}

template <class Real, bool VR>
cudaError_t launch_impl(const GridView<Real> &g, int v_begin, long long n_total, RayList<Real> rl,
                        ListView<Real> out, int *overflow_flag, cudaStream_t s) {
  if (n_total <= 0) return cudaSuccess;
  if (2 * (g.n_rb + g.n_sb) + 2 >= (1 << SLOT_BITS)) return cudaErrorInvalidValue;
  if (VR && n_total >= (1LL << 32)) return cudaErrorInvalidValue;   // ray_scalars divides in 32 bits
  const int threads = 128, warps = threads / 32;
  long long blocks = (n_total + warps - 1) / warps;
  const long long max_blocks = (long long) NUM_SMS * 16;
  if (blocks > max_blocks) blocks = max_blocks;
  static const bool force_general = getenv("B200RT_TRAVERSE_GENERAL") != nullptr;   // development aid: the exact ranking for every ray
  if (!g.pp && g.n_rb <= 128 && !force_general) {
    switch ((g.n_rb + 31) / 32) {
      case 1: return launch_fast<Real, VR, 1>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) blocks, threads, s);
      case 2: return launch_fast<Real, VR, 2>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) blocks, threads, s);
      case 3: return launch_fast<Real, VR, 3>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) blocks, threads, s);
      default: return launch_fast<Real, VR, 4>(g, v_begin, n_total, rl, out, overflow_flag, (unsigned) blocks, threads, s);
    }
  }
  const size_t smem = general_bytes_host(g) * warps;
  cudaError_t e = cudaFuncSetAttribute(traverse_kernel<Real, VR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  traverse_kernel<Real, VR><<<(unsigned) blocks, threads, smem, s>>>(g, v_begin, n_total, rl, out, overflow_flag);
  return cudaGetLastError();
}

} // namespace

template <class Real>
cudaError_t launch_sphere_table(const GridView<Real> &g, cudaStream_t s) {
  if (!g.sph_hdr || g.pp || g.n_rb > 128) return cudaErrorInvalidValue;
  const int threads = 128, warps = threads / 32, n_pairs = (g.n_rb - 1) * g.n_cls;
  const unsigned blocks = (unsigned) ((n_pairs + warps - 1) / warps);
  const size_t smem = (size_t) warps * 2 * g.n_rb * (sizeof(Real) + sizeof(int));
  int *hdr = const_cast<int *>(g.sph_hdr);
  Real *tde = const_cast<Real *>(g.sph_de), *td = const_cast<Real *>(g.sph_d);
  int *ti = const_cast<int *>(g.sph_i);
  switch ((g.n_rb + 31) / 32) {
    case 1: sphere_table_kernel<Real, 1><<<blocks, threads, smem, s>>>(g, hdr, tde, td, ti); break;
    case 2: sphere_table_kernel<Real, 2><<<blocks, threads, smem, s>>>(g, hdr, tde, td, ti); break;
    case 3: sphere_table_kernel<Real, 3><<<blocks, threads, smem, s>>>(g, hdr, tde, td, ti); break;
    default: sphere_table_kernel<Real, 4><<<blocks, threads, smem, s>>>(g, hdr, tde, td, ti); break;
  }
  return cudaGetLastError();
}
template cudaError_t launch_sphere_table<double>(const GridView<double> &, cudaStream_t);
template cudaError_t launch_sphere_table<float>(const GridView<float> &, cudaStream_t);

template <class Real>
cudaError_t launch_traverse_voxel_rays(const GridView<Real> &g, int v_begin, int v_end, ListView<Real> out,
                                       int *overflow_flag, cudaStream_t s) {
  RayList<Real> none = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return launch_impl<Real, true>(g, v_begin, (long long) (v_end - v_begin) * g.n_rays, none, out, overflow_flag, s);
}
template <class Real>
cudaError_t launch_traverse_list(const GridView<Real> &g, RayList<Real> rays, long long n, ListView<Real> out,
                                 int *overflow_flag, cudaStream_t s) {
  return launch_impl<Real, false>(g, 0, n, rays, out, overflow_flag, s);
}

template cudaError_t launch_traverse_voxel_rays<double>(const GridView<double> &, int, int, ListView<double>, int *, cudaStream_t);
template cudaError_t launch_traverse_voxel_rays<float>(const GridView<float> &, int, int, ListView<float>, int *, cudaStream_t);
template cudaError_t launch_traverse_list<double>(const GridView<double> &, RayList<double>, long long, ListView<double>, int *, cudaStream_t);
template cudaError_t launch_traverse_list<float>(const GridView<float> &, RayList<float>, long long, ListView<float>, int *, cudaStream_t);

} // namespace b200rt
