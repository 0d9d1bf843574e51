// b200rt_api.cu -- the C ABI (include/b200rt.h): context, uploads, orchestration of the
// traversal / influence / solve / brightness kernels.  Host code only.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <new>
#include <utility>
#include <vector>
#include "common.hpp"

using namespace b200rt;

// a context made by b200rt_create_multi owns no device state itself: every call fans out to its members (device_group.cu)
#define GROUP_DISPATCH(c, call) do { if ((c)->group) return call; } while (0)

namespace {

// Boundary-list scratch per batch: 8 GiB of the 180 GB, so that the bench workload (2.24e6 voxel rays, 1e6 lines of
// sight on the 100x60 grid: 7.0 and 3.1 GB of lists) runs as ONE batch per phase.  Every batch boundary drains the
// persistent brightness kernel (a single line of sight takes ~0.5 ms): measured 1.45 ms per boundary (r01n launch
// list: 686k LOS in 29.63 ms, 314k in 14.35 ms).  B200RT_SCRATCH_BYTES overrides it (tests force several batches).
constexpr size_t SCRATCH_BUDGET_BYTES = size_t(1) << 33;

struct PhaseTimer {
  b200rt_ctx *c;
  int phase;
  cudaEvent_t a, b;
  bool stopped = false;
  PhaseTimer(b200rt_ctx *ctx, int ph) : c(ctx), phase(ph) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, c->stream);
  }
  PhaseTimer(const PhaseTimer &) = delete;
  ~PhaseTimer() {   // an error path returned before stop(): the events are not handed to pending()
    if (!stopped) { cudaEventDestroy(a); cudaEventDestroy(b); }
  }
  void stop(int launches) {
    cudaEventRecord(b, c->stream);
    pending().push_back({phase, launches, a, b});
    stopped = true;
  }
  struct Rec { int phase, launches; cudaEvent_t a, b; };
  static std::vector<Rec> &pending() { static thread_local std::vector<Rec> v; return v; }
  static void reset(b200rt_ctx *c) {
    for (int p = 0; p < PH_COUNT; p++) { c->phase_ms[p] = 0; c->phase_launches[p] = 0; }
    for (auto &r : pending()) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    pending().clear();
  }
  static void collect(b200rt_ctx *c) {   // call after the stream has been synchronised
    for (auto &r : pending()) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) c->phase_ms[r.phase] += ms;
      c->phase_launches[r.phase] += r.launches;
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    pending().clear();
  }
};

// B200RT_DEBUG_SYNC=1: synchronise after every launch so a fault is attributed to its kernel
int dbg_sync(b200rt_ctx *c, const char *what) {
  static const bool on = getenv("B200RT_DEBUG_SYNC") != nullptr;
  if (!on) return B200RT_OK;
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e != cudaSuccess) return fail(c, B200RT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  return B200RT_OK;
}
#define DBG(c, what) do { if (int rc__ = dbg_sync(c, what)) return rc__; } while (0)

template <class Real>
GridView<Real> &gv(b200rt_ctx *c) { return *static_cast<GridView<Real> *>(c->grid_view); }

template <class Real>
EmissionView<Real> em_view(b200rt_ctx *c, int e) {
  const Emission &E = c->em[e];
  const int n = c->hg.n_vox;
  const Real *t = E.tabs.as<Real>();
  EmissionView<Real> v;
  v.T_ratio = t + 0 * n; v.density = t + 1 * n; v.dtau_species = t + 2 * n; v.dtau_absorber = t + 3 * n;
  v.T_ratio_pt = t + 4 * n; v.density_pt = t + 5 * n; v.dtau_species_pt = t + 6 * n; v.dtau_absorber_pt = t + 7 * n;
  v.phi = E.phi.as<Real>();
  v.mrec = E.mrec.as<Real>();
  v.sourcefn = E.S_real.as<Real>();
  v.rec_pt = E.rec_pt.as<Real>(); v.rec_avg = E.rec_avg.as<Real>();
  v.branching = (Real) E.branching; v.sigma_ref = (Real) E.sigma_ref; v.g_factor = (Real) E.g_factor;
  return v;
}

template <class Real>
int ensure_lists(b200rt_ctx *c, long long n_rays, ListView<Real> *lv) {
  const int cap = c->hg.cap;
  B200RT_CUDA(c, c->list_dist.ensure((size_t) n_rays * cap * sizeof(Real)));
  B200RT_CUDA(c, c->list_ent.ensure((size_t) n_rays * cap * sizeof(int)));
  B200RT_CUDA(c, c->list_len.ensure((size_t) n_rays * sizeof(int)));
  B200RT_CUDA(c, c->list_flag.ensure((size_t) n_rays * sizeof(int)));
  lv->dist = c->list_dist.as<Real>(); lv->ent = c->list_ent.as<int>();
  lv->len = c->list_len.as<int>(); lv->flag = c->list_flag.as<int>(); lv->cap = cap;
  return B200RT_OK;
}

// The longest-first order only pays when the queue is several times deeper than the machine (148 SMs x 4 CTAs x 32
// 4-lane groups = 18944 lines of sight in flight): below that every group gets at most a few lines of sight and the
// order only concentrates the long ones in the first CTAs.  B200RT_LOS_ORDER_MIN overrides (tests, diagnosis).
long long los_order_min() {
  if (const char *env = getenv("B200RT_LOS_ORDER_MIN")) return atoll(env);
  return 4LL * NUM_SMS * 4 * 32;
}

long long batch_capacity(b200rt_ctx *c, size_t real_bytes) {
  const size_t per_ray = (size_t) c->hg.cap * (real_bytes + sizeof(int)) + 2 * sizeof(int);
  size_t budget = SCRATCH_BUDGET_BYTES;
  if (const char *env = getenv("B200RT_SCRATCH_BYTES")) budget = (size_t) std::max(1LL, atoll(env));
  long long n = (long long) (budget / per_ray);
  return std::max<long long>(n, 1);
}

int check_overflow(b200rt_ctx *c) {
  B200RT_CUDA(c, c->host_words.ensure(4 * sizeof(unsigned long long)));
  int *flag_p = c->host_words.as<int>();
  *flag_p = 0;
  B200RT_CUDA(c, cudaMemcpyAsync(flag_p, c->work_counter.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  const int flag = *flag_p;
  if (flag)
    return fail(c, B200RT_ERR_CAPACITY, "a ray crossed more than 2*n_rb+n_sb boundaries (crossing list capacity)");
  return B200RT_OK;
}

// upload a host double array into a device Real array (widening/narrowing on the device)
template <class Real>
int upload_real(b200rt_ctx *c, const double *src, Real *dst, size_t n, DevBuf &stage);
template <>
int upload_real<double>(b200rt_ctx *c, const double *src, double *dst, size_t n, DevBuf &) {
  B200RT_CUDA(c, cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return B200RT_OK;
}
template <>
int upload_real<float>(b200rt_ctx *c, const double *src, float *dst, size_t n, DevBuf &stage) {
  B200RT_CUDA(c, stage.ensure(n * sizeof(double)));
  B200RT_CUDA(c, cudaMemcpyAsync(stage.p, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  B200RT_CUDA(c, launch_convert<float>(stage.as<double>(), dst, (long long) n, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));   // stage is reused by the caller
  return B200RT_OK;
}

// ------------------------------------------------------------------ influence
// ranges: the source-voxel ranges [begin, end) this call builds rows for (one range for a single GPU or a contiguous
// shard; several for the interleaved shards that balance the cost of low- and high-altitude rows across ranks)
template <class Real>
int influence_impl(b200rt_ctx *c, const std::vector<std::pair<int, int>> &ranges) {
  GridView<Real> &g = gv<Real>(c);
  const int n_vox = g.n_vox;
  PhaseTimer::reset(c);
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  B200RT_CUDA(c, cudaMemsetAsync(c->step_counter.p, 0, sizeof(unsigned long long), c->stream));
  int n_rows = 0;
  for (auto &r : ranges) n_rows += r.second - r.first;
  for (int e = 0; e < c->n_em; e++) {
    Emission &E = c->em[e];
    if (!E.defined) return fail(c, B200RT_ERR_STATE, "emission not defined");
    for (auto &r : ranges)
      if (r.second > r.first)
        B200RT_CUDA(c, cudaMemsetAsync(E.K.as<double>() + (size_t) r.first * n_vox, 0,
                                       (size_t) (r.second - r.first) * n_vox * sizeof(double), c->stream));
  }
  // Local slots 0 .. n_rows-1 run over the ranges in order.  One range: slot i is voxel first + i.  Several ranges
  // (interleaved multi-GPU shards): the slot -> voxel map goes to the device, so that a batch -- one traversal launch
  // + one march launch -- spans shard boundaries and the launch count does not grow with the number of shards.
  const bool mapped = ranges.size() > 1;
  std::vector<int> vox_of;
  if (mapped) {
    vox_of.reserve(n_rows);
    for (auto &r : ranges) for (int v = r.first; v < r.second; v++) vox_of.push_back(v);
    B200RT_CUDA(c, c->vox_map.ensure((size_t) n_rows * sizeof(int)));
    B200RT_CUDA(c, cudaMemcpyAsync(c->vox_map.p, vox_of.data(), (size_t) n_rows * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  const int first_voxel = ranges.empty() ? 0 : ranges[0].first;
  const long long cap_rays = batch_capacity(c, sizeof(Real));
  int vox_per_batch = (int) std::max<long long>(1, std::min<long long>(cap_rays / g.n_rays, std::max(n_rows, 1)));
  bool pushing = false;
  for (int e = 0; e < c->n_em; e++) {
    if (c->row_sink[e] && c->row_sink_n_vox[e] != n_vox)
      return fail(c, B200RT_ERR_STATE, "row sink was named for a different grid (b200rt_set_row_sink after the grid is set)");
    pushing = pushing || c->row_sink[e] != nullptr;
  }
  if (pushing && n_rows > 0) {     // several batches, so that the DMA of one overlaps the march of the next
    if (const char *env = getenv("B200RT_ROW_PUSH_BATCHES")) c->row_push_batches = std::max(1, atoi(env));
    vox_per_batch = std::max(1, std::min(vox_per_batch, (n_rows + c->row_push_batches - 1) / c->row_push_batches));
    if (!c->copy_stream) B200RT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->ev_rows) B200RT_CUDA(c, cudaEventCreateWithFlags(&c->ev_rows, cudaEventDisableTiming));
  }
  ListView<Real> lv;
  const long long need = std::max<long long>((long long) vox_per_batch * g.n_rays, n_vox);
  if (int rc = ensure_lists<Real>(c, need, &lv)) return rc;
  int *overflow = c->work_counter.as<int>() + 1;

  for (int lb = 0; lb < n_rows; lb += vox_per_batch) {
    const int le = std::min(n_rows, lb + vox_per_batch);
    GridView<Real> gb = g;
    int vb = first_voxel + lb, ve = first_voxel + le;      // unmapped: the voxel range itself
    if (mapped) { gb.vox_map = c->vox_map.as<int>() + lb; vb = 0; ve = le - lb; }
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_voxel_rays<Real>(gb, vb, ve, lv, overflow, c->stream));
      t.stop(1);
      DBG(c, "traverse_voxel_rays");
    }
    for (int e = 0; e < c->n_em; e++) {
      PhaseTimer t(c, PH_INFLUENCE);
      B200RT_CUDA(c, launch_influence<Real>(gb, em_view<Real>(c, e), vb, ve, lv, c->em[e].K.as<double>(),
                                            c->work_counter.as<int>(),
                                            e == 0 ? c->step_counter.as<unsigned long long>() : nullptr, c->stream));
      t.stop(1);
      DBG(c, "influence march");
    }
    if (pushing) {   // the rows of this batch are final: hand them to the solving GPU (peer memory, copy engine, NVLink)
      B200RT_CUDA(c, cudaEventRecord(c->ev_rows, c->stream));
      B200RT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_rows, 0));
      int run_lo = lb;                                      // contiguous runs of voxels inside the batch
      while (run_lo < le) {
        int run_hi = run_lo + 1;
        const int v_lo = mapped ? vox_of[run_lo] : first_voxel + run_lo;
        while (run_hi < le && (mapped ? vox_of[run_hi] : first_voxel + run_hi) == v_lo + (run_hi - run_lo)) run_hi++;
        for (int e = 0; e < c->n_em; e++)
          if (c->row_sink[e])
            B200RT_CUDA(c, cudaMemcpyAsync(static_cast<double *>(c->row_sink[e]) + (size_t) v_lo * n_vox,
                                           c->em[e].K.as<double>() + (size_t) v_lo * n_vox,
                                           (size_t) (run_hi - run_lo) * n_vox * sizeof(double), cudaMemcpyDeviceToDevice,
                                           c->copy_stream));
        run_lo = run_hi;
      }
    }
  }
  // single scattering: one sun-ward ray per voxel (every rank computes all of them: n_vox rays)
  {
    const Real *sp = c->sun_rays.as<Real>();
    RayList<Real> rl;
    rl.r = sp + 0 * (size_t) n_vox; rl.z = sp + 1 * (size_t) n_vox; rl.t = sp + 2 * (size_t) n_vox;
    rl.cost = sp + 3 * (size_t) n_vox; rl.lz = sp + 4 * (size_t) n_vox;
    const int *ip = reinterpret_cast<const int *>(sp + 5 * (size_t) n_vox);
    rl.i_voxel = ip;
    const int *shadow = ip + n_vox;
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_list<Real>(g, rl, n_vox, lv, overflow, c->stream));
      t.stop(1);
      DBG(c, "traverse sun rays");
    }
    for (int e = 0; e < c->n_em; e++) {
      PhaseTimer t(c, PH_INFLUENCE);
      Emission &E = c->em[e];
      B200RT_CUDA(c, launch_single_scattering<Real>(g, em_view<Real>(c, e), lv, shadow, E.S0.as<double>(),
                                                    E.tau_sp.as<double>(), E.tau_abs.as<double>(),
                                                    c->work_counter.as<int>(), c->stream));
      t.stop(1);
      DBG(c, "single scattering march");
    }
  }
  B200RT_CUDA(c, c->host_words.ensure(4 * sizeof(unsigned long long)));
  unsigned long long *steps_p = c->host_words.as<unsigned long long>() + 1;
  B200RT_CUDA(c, cudaMemcpyAsync(steps_p, c->step_counter.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  const int rc_overflow = check_overflow(c);   // synchronises
  const unsigned long long steps = *steps_p;
  if (pushing) B200RT_CUDA(c, cudaStreamSynchronize(c->copy_stream));   // the rows have landed on the solving GPU
  if (rc_overflow) return rc_overflow;
  PhaseTimer::collect(c);
  c->last_steps = (long long) steps;
  for (int e = 0; e < c->n_em; e++) { c->em[e].have_K = true; c->em[e].have_S = false; }
  return B200RT_OK;
}

int solve_impl(b200rt_ctx *c, bool reset_timer) {
  const int n = c->hg.n_vox;
  if (reset_timer) PhaseTimer::reset(c);
  for (int e = 0; e < c->n_em; e++) {
    Emission &E = c->em[e];
    if (!E.have_K) return fail(c, B200RT_ERR_STATE, "b200rt_solve: influence matrix not built");
    PhaseTimer t(c, PH_SOLVE);
    SolveResult r = {0, 0, 0};
    if (int rc = solve_dense(c, n, E.K.as<double>(), E.branching, E.S0.as<double>(), E.S.as<double>(), &r)) return rc;
    t.stop(r.launches);
    E.residual = r.residual;
    if (c->precision == B200RT_F64)
      B200RT_CUDA(c, launch_convert<double>(E.S.as<double>(), E.S_real.as<double>(), n, c->stream));
    else
      B200RT_CUDA(c, launch_convert<float>(E.S.as<double>(), E.S_real.as<float>(), n, c->stream));
    E.have_S = true;
    E.rec_dirty = true;
  }
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  PhaseTimer::collect(c);
  return B200RT_OK;
}

// ------------------------------------------------------------------ brightness
// With `io` (double builds only) the lines of sight come from, and the results go to, HOST arrays.  When the set needs
// several batches (more lists than the scratch budget holds) they are pipelined -- batch b+1's nine input slices travel
// on copy_stream and batch b-1's result slices on out_stream while batch b is traversed and marched on the compute
// stream -- so that only the first upload and the last download are exposed (b200rt_brightness; with pageable host
// memory the copies degrade to staged ones, still correct).  Batches are NOT made smaller to get more overlap: a
// forced 4-way split of the 1e6-LOS bench set cost more in kernel tails (+3 ms) than the hidden copies saved (2 ms).
struct HostLos {
  int n;
  const double *const *src;   // [9]
  double *const *dst;         // [4], entries may be null
  long long out_stride;       // the caller's result arrays are [n_emissions][out_stride]; this call fills
  long long out_offset;       // [out_offset, out_offset + n) of each row (a device group hands every member a slice)
};

template <class Real>
int brightness_impl(b200rt_ctx *c, int n_subsamples, const HostLos *io = nullptr) {
  if (n_subsamples == 1 || n_subsamples < 0)
    return fail(c, B200RT_ERR_ARG, "n_subsamples must be 0 or > 1 (RT_grid.hpp:237)");
  if (io) {
    B200RT_CUDA(c, c->los_in.ensure((size_t) 9 * io->n * sizeof(Real)));
    c->n_los = io->n;
    c->los_done = false;
    if (!c->copy_stream) B200RT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->out_stream) B200RT_CUDA(c, cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
  }
  if (c->n_los <= 0) return fail(c, B200RT_ERR_STATE, "no lines of sight uploaded");
  if (c->hg.pp)
    return fail(c, B200RT_ERR_STATE, "interp_weights not implemented in grid_plane_parallel (grid_plane_parallel.hpp:304-311)");
  for (int e = 0; e < c->n_em; e++)
    if (!c->em[e].have_S) return fail(c, B200RT_ERR_STATE, "source function not available (solve or set_sourcefn first)");
  GridView<Real> &g = gv<Real>(c);
  PhaseTimer::reset(c);
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  B200RT_CUDA(c, cudaMemsetAsync(c->step_counter.p, 0, sizeof(unsigned long long), c->stream));
  const long long n = c->n_los;
  const long long per_batch = std::min<long long>(batch_capacity(c, sizeof(Real)), n);
  ListView<Real> lv;
  if (int rc = ensure_lists<Real>(c, per_batch, &lv)) return rc;
  B200RT_CUDA(c, c->los_out.ensure((size_t) c->n_em * 4 * n * sizeof(Real)));
  B200RT_CUDA(c, c->los_order.ensure(((size_t) per_batch + 2 * (size_t) (c->hg.cap + 1)) * sizeof(int)));
  const Real *li = c->los_in.as<Real>();
  // Results for PAGEABLE caller arrays are downloaded into page-locked scratch and copied out after the stream has
  // drained: a device-to-host copy into pageable memory blocks inside the runtime until the kernels before it have
  // finished, holding a lock that stalls every other context of the process (the sweep's contexts ran their brightness
  // calls one after the other because of it).  Page-locked caller arrays (the bench's) are written directly.
  double *stage_out = nullptr;
  if (io) {
    bool pageable = false;
    for (int q = 0; q < 4 && !pageable; q++)
      if (io->dst[q]) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, io->dst[q]) != cudaSuccess) { cudaGetLastError(); pageable = true; }
        else pageable = (at.type == cudaMemoryTypeUnregistered);
      }
    if (pageable) {
      B200RT_CUDA(c, c->host_out.ensure((size_t) c->n_em * 4 * n * sizeof(double)));
      stage_out = c->host_out.as<double>();
    }
  }
  std::vector<cudaEvent_t> io_events;
  struct EventGuard {
    std::vector<cudaEvent_t> &v;
    ~EventGuard() { for (auto e : v) cudaEventDestroy(e); }
  } io_guard{io_events};
  auto io_event = [&](cudaStream_t on, cudaEvent_t *out) -> cudaError_t {   // event recorded on `on`; destroyed at return
    cudaEvent_t e;
    cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (rc != cudaSuccess) return rc;
    io_events.push_back(e);
    if (out) *out = e;
    return cudaEventRecord(e, on);
  };
  // the traversal reads r, z, t, cos(theta), line_z (arrays 3, 2, 4, 8, 7); x, y, line_x, line_y are only read by the
  // march, so they travel while the batch is being traversed
  cudaEvent_t uploaded_trav = nullptr, uploaded_all = nullptr;              // of the latest upload_batch
  auto upload_batch = [&](long long first) -> cudaError_t {
    const long long count = std::min(per_batch, n - first);
    static const int order_of_arrays[9] = {3, 2, 4, 8, 7, 0, 1, 5, 6};
    for (int k = 0; k < 9; k++) {
      const int a = order_of_arrays[k];
      cudaError_t rc = cudaMemcpyAsync(c->los_in.as<double>() + (size_t) a * n + first, io->src[a] + first,
                                       (size_t) count * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream);
      if (rc != cudaSuccess) return rc;
      if (k == 4) {
        rc = io_event(c->copy_stream, &uploaded_trav);
        if (rc != cudaSuccess) return rc;
      }
    }
    return io_event(c->copy_stream, &uploaded_all);
  };
  EmissionView<Real> ev[MAX_EMISSIONS];
  for (int e = 0; e < c->n_em; e++) {
    ev[e] = em_view<Real>(c, e);
    Emission &E = c->em[e];
    if (E.rec_dirty) {
      B200RT_CUDA(c, launch_pack_records<Real>(ev[e], g.n_vox, E.rec_pt.as<Real>(), E.rec_avg.as<Real>(), c->stream));
      E.rec_dirty = false;
    }
  }
  int *overflow = c->work_counter.as<int>() + 1;
  for (long long first = 0; first < n; first += per_batch) {
    const long long count = std::min(per_batch, n - first);
    RayList<Real> rl;
    rl.r = li + 3 * n + first; rl.z = li + 2 * n + first; rl.t = li + 4 * n + first;
    rl.cost = li + 8 * n + first; rl.lz = li + 7 * n + first; rl.i_voxel = nullptr;
    if (io) {
      if (first == 0) B200RT_CUDA(c, upload_batch(0));
      B200RT_CUDA(c, cudaStreamWaitEvent(c->stream, uploaded_trav, 0));   // this batch's traversal slices have arrived
    }
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_list<Real>(g, rl, count, lv, overflow, c->stream));
      t.stop(1);
    }
    if (io) B200RT_CUDA(c, cudaStreamWaitEvent(c->stream, uploaded_all, 0));
    const int *order = nullptr;
    if (lv.cap <= LOS_ORDER_MAX_CAP && count >= los_order_min()) {
      PhaseTimer t(c, PH_ORDER);   // histogram, prefix, scatter: timed and counted apart from the march they feed
      int *bins = c->los_order.as<int>(), *ord = bins + 2 * (lv.cap + 1);
      B200RT_CUDA(c, launch_los_order(lv.len, count, lv.cap, bins, ord, c->stream));
      order = ord;
      t.stop(3);
    }
    {
      PhaseTimer t(c, PH_BRIGHTNESS);
      B200RT_CUDA(c, launch_brightness<Real>(g, ev, c->n_em, li, n, first, count, lv, n_subsamples,
                                             c->los_out.as<Real>(), n, c->work_counter.as<int>(),
                                             c->step_counter.as<unsigned long long>(), order, c->stream));
      t.stop(1);
    }
    if (io) {
      cudaEvent_t done;
      B200RT_CUDA(c, io_event(c->stream, &done));
      // order matters for pageable host memory, whose copies block the host: the kernels of this batch are queued
      // first, the next batch's upload runs beside them, and only then does the download wait for them
      if (first + count < n) B200RT_CUDA(c, upload_batch(first + count));
      B200RT_CUDA(c, cudaStreamWaitEvent(c->out_stream, done, 0));
      for (int e = 0; e < c->n_em; e++)
        for (int q = 0; q < 4; q++)
          if (io->dst[q])
            B200RT_CUDA(c, cudaMemcpyAsync(stage_out ? stage_out + ((size_t) e * 4 + q) * n + first
                                                     : io->dst[q] + (size_t) e * io->out_stride + io->out_offset + first,
                                           c->los_out.as<double>() + ((size_t) e * 4 + q) * n + first,
                                           (size_t) count * sizeof(double), cudaMemcpyDeviceToHost, c->out_stream));
    }
  }
  B200RT_CUDA(c, c->host_words.ensure(4 * sizeof(unsigned long long)));
  unsigned long long *substeps_p = c->host_words.as<unsigned long long>() + 1;
  B200RT_CUDA(c, cudaMemcpyAsync(substeps_p, c->step_counter.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  const int rc_overflow = check_overflow(c);
  const unsigned long long substeps = *substeps_p;
  if (io) B200RT_CUDA(c, cudaStreamSynchronize(c->out_stream));   // nothing is in flight into the caller's arrays at return
  if (rc_overflow) return rc_overflow;
  if (stage_out)
    for (int e = 0; e < c->n_em; e++)
      for (int q = 0; q < 4; q++)
        if (io->dst[q])
          std::memcpy(io->dst[q] + (size_t) e * io->out_stride + io->out_offset, stage_out + ((size_t) e * 4 + q) * n,
                      (size_t) n * sizeof(double));
  PhaseTimer::collect(c);
  c->last_substeps = (long long) substeps;
  c->los_done = true;
  return B200RT_OK;
}

template <class Real>
int los_upload_impl(b200rt_ctx *c, int n, const double *const src[9]) {
  B200RT_CUDA(c, c->los_in.ensure((size_t) 9 * n * sizeof(Real)));
  DevBuf stage;
  int rc = B200RT_OK;
  for (int a = 0; a < 9 && rc == B200RT_OK; a++)
    rc = upload_real<Real>(c, src[a], c->los_in.as<Real>() + (size_t) a * n, n, stage);
  if (rc == B200RT_OK) {
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(c, B200RT_ERR_CUDA, cudaGetErrorString(e));
  }
  stage.release();
  c->n_los = n;
  c->los_done = false;
  return rc;
}

template <class Real>
int los_download_impl(b200rt_ctx *c, double *const dst[4], long long stride, long long offset) {
  const long long n = c->n_los;
  const Real *o = c->los_out.as<Real>();
  DevBuf stage;
  for (int e = 0; e < c->n_em; e++)
    for (int q = 0; q < 4; q++) {
      if (!dst[q]) continue;
      const Real *src = o + ((size_t) e * 4 + q) * n;
      double *out = dst[q] + (size_t) e * stride + offset;
      if (sizeof(Real) == sizeof(double)) {
        B200RT_CUDA(c, cudaMemcpyAsync(out, src, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      } else {
        std::vector<float> tmp(n);
        B200RT_CUDA(c, cudaMemcpyAsync(tmp.data(), src, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
        for (long long i = 0; i < n; i++) out[i] = tmp[i];
      }
    }
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  stage.release();
  return B200RT_OK;
}

// compact fixed-stride device lists into the caller's concatenated arrays
template <class Real>
int fetch_lists(b200rt_ctx *c, const ListView<Real> &lv, long long n_rays, long long capacity, long long *pos,
                int *len, int *exits_bottom, int *entering, double *distance) {
  const int cap = lv.cap;
  std::vector<int> hl(n_rays), hf(n_rays), he((size_t) n_rays * cap);
  std::vector<Real> hd((size_t) n_rays * cap);
  B200RT_CUDA(c, cudaMemcpyAsync(hl.data(), lv.len, n_rays * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaMemcpyAsync(hf.data(), lv.flag, n_rays * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaMemcpyAsync(he.data(), lv.ent, (size_t) n_rays * cap * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaMemcpyAsync(hd.data(), lv.dist, (size_t) n_rays * cap * sizeof(Real), cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (long long i = 0; i < n_rays; i++) {
    len[i] = hl[i];
    exits_bottom[i] = hf[i] & 1;
    if (*pos + hl[i] > capacity) return fail(c, B200RT_ERR_ARG, "output capacity too small for the boundary lists");
    for (int k = 0; k < hl[i]; k++) {
      entering[*pos + k] = he[(size_t) i * cap + k];
      distance[*pos + k] = (double) hd[(size_t) i * cap + k];
    }
    *pos += hl[i];
  }
  return B200RT_OK;
}

template <class Real>
int traverse_voxel_rays_impl(b200rt_ctx *c, int v_begin, int v_end, long long capacity, int *len, int *exits_bottom,
                             int *entering, double *distance, long long *n_entries) {
  GridView<Real> &g = gv<Real>(c);
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  const long long cap_rays = std::min<long long>(batch_capacity(c, sizeof(Real)), 1 << 20);
  const int vpb = (int) std::max<long long>(1, std::min<long long>(cap_rays / g.n_rays, v_end - v_begin));
  ListView<Real> lv;
  if (int rc = ensure_lists<Real>(c, (long long) vpb * g.n_rays, &lv)) return rc;
  long long pos = 0;
  for (int vb = v_begin; vb < v_end; vb += vpb) {
    const int ve = std::min(v_end, vb + vpb);
    B200RT_CUDA(c, launch_traverse_voxel_rays<Real>(g, vb, ve, lv, c->work_counter.as<int>() + 1, c->stream));
    const long long nr = (long long) (ve - vb) * g.n_rays;
    const long long off = (long long) (vb - v_begin) * g.n_rays;
    if (int rc = fetch_lists<Real>(c, lv, nr, capacity, &pos, len + off, exits_bottom + off, entering, distance)) return rc;
  }
  if (n_entries) *n_entries = pos;
  return check_overflow(c);
}

template <class Real>
int traverse_los_impl(b200rt_ctx *c, long long capacity, int *len, int *exits_bottom, int *entering,
                      double *distance, long long *n_entries) {
  GridView<Real> &g = gv<Real>(c);
  if (c->n_los <= 0) return fail(c, B200RT_ERR_STATE, "no lines of sight uploaded");
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  const long long n = c->n_los;
  const long long per_batch = std::min<long long>(std::min<long long>(batch_capacity(c, sizeof(Real)), 1 << 18), n);
  ListView<Real> lv;
  if (int rc = ensure_lists<Real>(c, per_batch, &lv)) return rc;
  const Real *li = c->los_in.as<Real>();
  long long pos = 0;
  for (long long first = 0; first < n; first += per_batch) {
    const long long count = std::min(per_batch, n - first);
    RayList<Real> rl;
    rl.r = li + 3 * n + first; rl.z = li + 2 * n + first; rl.t = li + 4 * n + first;
    rl.cost = li + 8 * n + first; rl.lz = li + 7 * n + first; rl.i_voxel = nullptr;
    B200RT_CUDA(c, launch_traverse_list<Real>(g, rl, count, lv, c->work_counter.as<int>() + 1, c->stream));
    if (int rc = fetch_lists<Real>(c, lv, count, capacity, &pos, len + first, exits_bottom + first, entering, distance)) return rc;
  }
  if (n_entries) *n_entries = pos;
  return check_overflow(c);
}

template <class Real>
int set_singlet_impl(b200rt_ctx *c, int e, const double *const arr[8]) {
  const int n = c->hg.n_vox;
  Emission &E = c->em[e];
  B200RT_CUDA(c, E.tabs.ensure((size_t) 8 * n * sizeof(Real)));
  B200RT_CUDA(c, E.phi.ensure((size_t) n * N_LAMBDA * sizeof(Real)));
  B200RT_CUDA(c, E.mrec.ensure((size_t) n * 2 * N_LAMBDA * sizeof(Real)));
  B200RT_CUDA(c, E.K.ensure((size_t) n * n * sizeof(double)));
  B200RT_CUDA(c, E.S0.ensure(n * sizeof(double)));
  B200RT_CUDA(c, E.tau_sp.ensure(n * sizeof(double)));
  B200RT_CUDA(c, E.tau_abs.ensure(n * sizeof(double)));
  B200RT_CUDA(c, E.S.ensure(n * sizeof(double)));
  B200RT_CUDA(c, E.S_real.ensure(n * sizeof(Real)));
  B200RT_CUDA(c, E.rec_pt.ensure((size_t) n * 8 * sizeof(Real)));
  B200RT_CUDA(c, E.rec_avg.ensure((size_t) n * 8 * sizeof(Real)));
  E.rec_dirty = true;
  DevBuf stage;
  int rc = B200RT_OK;
  for (int a = 0; a < 8 && rc == B200RT_OK; a++)
    rc = upload_real<Real>(c, arr[a], E.tabs.as<Real>() + (size_t) a * n, n, stage);
  if (rc == B200RT_OK) {
    cudaError_t er = launch_phi_table<Real>(E.tabs.as<Real>(), E.tabs.as<Real>() + 2 * (size_t) n, E.tabs.as<Real>() + 3 * (size_t) n, n,
                                            E.phi.as<Real>(), E.mrec.as<Real>(), c->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(c->stream);
    if (er != cudaSuccess) rc = fail(c, B200RT_ERR_CUDA, cudaGetErrorString(er));
  }
  stage.release();
  return rc;
}

bool is64(const b200rt_ctx *c) { return c->precision == B200RT_F64; }

// ------------------------------------------------------------------ multiplet emission
template <class Real>
MultView<Real> mult_view(b200rt_ctx *c) {
  Multiplet &M = c->mult;
  const size_t n = c->hg.n_vox;
  Real *t = M.tabs.as<Real>();
  MultView<Real> v;
  v.T = t; v.T_pt = t + n; v.nabs = t + 2 * n; v.nabs_pt = t + 3 * n;
  for (int l = 0; l < MULT_MAX_LOWER; l++) { v.n[l] = t + (4 + l) * n; v.n_pt[l] = t + (7 + l) * n; }
  v.rec_step = M.rec_step.as<Real>(); v.rec_org = M.rec_org.as<Real>(); v.rec_w0 = M.rec_w0.as<Real>();
  v.tsv = M.tsv.as<Real>(); v.tav = M.tav.as<Real>();
  v.S = M.S_real.as<Real>();
  v.rec_pt = M.rec_pt.as<Real>(); v.rec_avg = M.rec_avg.as<Real>();
  return v;
}

template <class Real>
int set_multiplet_impl(b200rt_ctx *c, const double *const arr[6]) {
  Multiplet &M = c->mult;
  const b200rt_multiplet_desc &d = M.d;
  const size_t n = c->hg.n_vox, ne = n * d.n_upper;
  const size_t nlp = (d.n_lambda + MULT_LPR - 1) / MULT_LPR, slots = n * MULT_LPR * nlp;
  B200RT_CUDA(c, M.tabs.ensure(10 * n * sizeof(Real)));
  B200RT_CUDA(c, cudaMemsetAsync(M.tabs.p, 0, 10 * n * sizeof(Real), c->stream));
  B200RT_CUDA(c, M.rec_step.ensure(slots * (d.n_multiplets + d.n_lines) * sizeof(Real)));
  B200RT_CUDA(c, M.rec_org.ensure(slots * d.n_lines * sizeof(Real)));
  B200RT_CUDA(c, M.rec_w0.ensure(slots * d.n_lines * sizeof(Real)));
  B200RT_CUDA(c, M.tsv.ensure(n * d.n_lines * sizeof(Real)));
  B200RT_CUDA(c, M.tav.ensure(n * d.n_lines * sizeof(Real)));
  B200RT_CUDA(c, M.K.ensure(ne * ne * sizeof(double)));
  B200RT_CUDA(c, M.S0.ensure(ne * sizeof(double)));
  B200RT_CUDA(c, M.S.ensure(ne * sizeof(double)));
  B200RT_CUDA(c, M.S_real.ensure(ne * sizeof(Real)));
  B200RT_CUDA(c, M.tau_sp.ensure(n * d.n_lines * sizeof(double)));
  B200RT_CUDA(c, M.tau_abs.ensure(n * d.n_lines * sizeof(double)));
  B200RT_CUDA(c, M.rec_pt.ensure(n * MULT_REC * sizeof(Real)));
  B200RT_CUDA(c, M.rec_avg.ensure(n * MULT_REC * sizeof(Real)));
  Real *t = M.tabs.as<Real>();
  DevBuf stage;
  int rc = B200RT_OK;
  // arr: species_density [n_lower][n], species_density_pt, T, T_pt, absorber, absorber_pt
  for (int l = 0; l < d.n_lower && rc == B200RT_OK; l++) {
    rc = upload_real<Real>(c, arr[0] + (size_t) l * n, t + (4 + l) * n, n, stage);
    if (rc == B200RT_OK) rc = upload_real<Real>(c, arr[1] + (size_t) l * n, t + (7 + l) * n, n, stage);
  }
  for (int a = 0; a < 4 && rc == B200RT_OK; a++) rc = upload_real<Real>(c, arr[2 + a], t + (size_t) a * n, n, stage);
  if (rc == B200RT_OK) {
    cudaError_t er = launch_mult_tables<Real>(d, mult_view<Real>(c), (int) n, c->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(c->stream);
    if (er != cudaSuccess) rc = fail(c, B200RT_ERR_CUDA, cudaGetErrorString(er));
  }
  stage.release();
  return rc;
}

template <class Real>
int mult_influence_impl(b200rt_ctx *c, int v_begin, int v_end) {
  GridView<Real> &g = gv<Real>(c);
  Multiplet &M = c->mult;
  const b200rt_multiplet_desc &d = M.d;
  const int n_vox = g.n_vox;
  const size_t ne = (size_t) n_vox * d.n_upper;
  PhaseTimer::reset(c);
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  B200RT_CUDA(c, cudaMemsetAsync(c->step_counter.p, 0, sizeof(unsigned long long), c->stream));
  if (v_end > v_begin)
    B200RT_CUDA(c, cudaMemsetAsync(M.K.as<double>() + (size_t) v_begin * d.n_upper * ne, 0,
                                   (size_t) (v_end - v_begin) * d.n_upper * ne * sizeof(double), c->stream));
  const long long cap_rays = batch_capacity(c, sizeof(Real));
  const int vox_per_batch = (int) std::max<long long>(1, std::min<long long>(cap_rays / g.n_rays, v_end - v_begin));
  ListView<Real> lv;
  const long long need = std::max<long long>((long long) vox_per_batch * g.n_rays, n_vox);
  if (int rc = ensure_lists<Real>(c, need, &lv)) return rc;
  int *overflow = c->work_counter.as<int>() + 1;
  MultView<Real> mv = mult_view<Real>(c);
  for (int vb = v_begin; vb < v_end; vb += vox_per_batch) {
    const int ve = std::min(v_end, vb + vox_per_batch);
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_voxel_rays<Real>(g, vb, ve, lv, overflow, c->stream));
      t.stop(1);
    }
    {
      PhaseTimer t(c, PH_INFLUENCE);
      B200RT_CUDA(c, launch_mult_influence<Real>(d, g, mv, vb, ve, lv, M.K.as<double>(), c->work_counter.as<int>(),
                                                 c->step_counter.as<unsigned long long>(), c->stream));
      t.stop(1);
      DBG(c, "multiplet influence march");
    }
  }
  {
    const Real *sp = c->sun_rays.as<Real>();
    RayList<Real> rl;
    rl.r = sp + 0 * (size_t) n_vox; rl.z = sp + 1 * (size_t) n_vox; rl.t = sp + 2 * (size_t) n_vox;
    rl.cost = sp + 3 * (size_t) n_vox; rl.lz = sp + 4 * (size_t) n_vox;
    const int *ip = reinterpret_cast<const int *>(sp + 5 * (size_t) n_vox);
    rl.i_voxel = ip;
    const int *shadow = ip + n_vox;
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_list<Real>(g, rl, n_vox, lv, overflow, c->stream));
      t.stop(1);
    }
    {
      PhaseTimer t(c, PH_INFLUENCE);
      B200RT_CUDA(c, cudaMemsetAsync(M.S0.p, 0, ne * sizeof(double), c->stream));
      B200RT_CUDA(c, launch_mult_single_scattering<Real>(d, g, mv, lv, shadow, M.S0.as<double>(), M.tau_sp.as<double>(),
                                                         M.tau_abs.as<double>(), c->work_counter.as<int>(), c->stream));
      t.stop(1);
      DBG(c, "multiplet single scattering");
    }
  }
  B200RT_CUDA(c, c->host_words.ensure(4 * sizeof(unsigned long long)));
  unsigned long long *steps_p = c->host_words.as<unsigned long long>() + 1;
  B200RT_CUDA(c, cudaMemcpyAsync(steps_p, c->step_counter.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  if (int rc = check_overflow(c)) return rc;
  const unsigned long long steps = *steps_p;
  PhaseTimer::collect(c);
  c->last_steps = (long long) steps;
  M.have_K = true; M.have_S = false;
  return B200RT_OK;
}

int mult_solve_impl(b200rt_ctx *c, bool reset_timer) {
  Multiplet &M = c->mult;
  const int ne = c->hg.n_vox * M.d.n_upper;
  if (reset_timer) PhaseTimer::reset(c);
  if (!M.have_K) return fail(c, B200RT_ERR_STATE, "b200rt_solve: influence matrix not built");
  PhaseTimer t(c, PH_SOLVE);
  SolveResult r = {0, 0, 0};
  // multiplet_CFR_emission::pre_solve is empty: kernel = I - K (multiplet_CFR_emission.hpp:408; emission_voxels.hpp:170-176)
  if (int rc = solve_dense(c, ne, M.K.as<double>(), 1.0, M.S0.as<double>(), M.S.as<double>(), &r)) return rc;
  t.stop(r.launches);
  M.residual = r.residual;
  if (c->precision == B200RT_F64) B200RT_CUDA(c, launch_convert<double>(M.S.as<double>(), M.S_real.as<double>(), ne, c->stream));
  else B200RT_CUDA(c, launch_convert<float>(M.S.as<double>(), M.S_real.as<float>(), ne, c->stream));
  M.have_S = true;
  M.rec_dirty = true;
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  PhaseTimer::collect(c);
  return B200RT_OK;
}

template <class Real>
int mult_brightness_impl(b200rt_ctx *c, int n_subsamples) {
  Multiplet &M = c->mult;
  const b200rt_multiplet_desc &d = M.d;
  if (n_subsamples == 1 || n_subsamples < 0)
    return fail(c, B200RT_ERR_ARG, "n_subsamples must be 0 or > 1 (RT_grid.hpp:237)");
  if (c->n_los <= 0) return fail(c, B200RT_ERR_STATE, "no lines of sight uploaded");
  if (c->hg.pp)
    return fail(c, B200RT_ERR_STATE, "interp_weights not implemented in grid_plane_parallel (grid_plane_parallel.hpp:304-311)");
  if (!M.have_S) return fail(c, B200RT_ERR_STATE, "source function not available (solve or set_sourcefn first)");
  GridView<Real> &g = gv<Real>(c);
  PhaseTimer::reset(c);
  B200RT_CUDA(c, cudaMemsetAsync(c->work_counter.p, 0, 2 * sizeof(int), c->stream));
  const long long n = c->n_los;
  const long long per_batch = std::min<long long>(batch_capacity(c, sizeof(Real)), n);
  ListView<Real> lv;
  if (int rc = ensure_lists<Real>(c, per_batch, &lv)) return rc;
  const size_t n_out = 3 * d.n_lines + d.n_lower;
  B200RT_CUDA(c, c->los_out.ensure(n_out * n * sizeof(Real)));
  B200RT_CUDA(c, c->los_order.ensure(((size_t) per_batch + 2 * (size_t) (c->hg.cap + 1)) * sizeof(int)));
  const Real *li = c->los_in.as<Real>();
  MultView<Real> mv = mult_view<Real>(c);
  if (M.rec_dirty) {
    B200RT_CUDA(c, launch_mult_pack<Real>(d, mv, g.n_vox, c->stream));
    M.rec_dirty = false;
  }
  int *overflow = c->work_counter.as<int>() + 1;
  for (long long first = 0; first < n; first += per_batch) {
    const long long count = std::min(per_batch, n - first);
    RayList<Real> rl;
    rl.r = li + 3 * n + first; rl.z = li + 2 * n + first; rl.t = li + 4 * n + first;
    rl.cost = li + 8 * n + first; rl.lz = li + 7 * n + first; rl.i_voxel = nullptr;
    {
      PhaseTimer t(c, PH_TRAVERSE);
      B200RT_CUDA(c, launch_traverse_list<Real>(g, rl, count, lv, overflow, c->stream));
      t.stop(1);
    }
    const int *order = nullptr;
    if (lv.cap <= LOS_ORDER_MAX_CAP && count >= los_order_min()) {
      PhaseTimer t(c, PH_ORDER);
      int *bins = c->los_order.as<int>(), *ord = bins + 2 * (lv.cap + 1);
      B200RT_CUDA(c, launch_los_order(lv.len, count, lv.cap, bins, ord, c->stream));
      order = ord;
      t.stop(3);
    }
    {
      PhaseTimer t(c, PH_BRIGHTNESS);
      B200RT_CUDA(c, launch_mult_brightness<Real>(d, g, mv, li, n, first, count, lv, n_subsamples, c->los_out.as<Real>(), n,
                                                  c->work_counter.as<int>(), order, c->stream));
      t.stop(1);
    }
  }
  if (int rc = check_overflow(c)) return rc;
  PhaseTimer::collect(c);
  c->los_done = true;
  return B200RT_OK;
}

// multiplet outputs [3 n_lines + n_lower][n_los] -> brightness, tau_species_final, tau_absorber_final [n_lines][n],
// species_col_dens [n_lower][n]
template <class Real>
int mult_los_download_impl(b200rt_ctx *c, double *const dst[4], long long stride, long long offset) {
  const long long n = c->n_los;
  const b200rt_multiplet_desc &d = c->mult.d;
  const Real *o = c->los_out.as<Real>();
  const int rows[4] = {d.n_lines, d.n_lines, d.n_lines, d.n_lower};
  size_t row0 = 0;
  std::vector<float> tmp;
  for (int q = 0; q < 4; q++) {
    if (dst[q])
      for (int r = 0; r < rows[q]; r++) {
        const Real *src = o + (row0 + r) * (size_t) n;
        double *out = dst[q] + (size_t) r * stride + offset;
        if (sizeof(Real) == sizeof(double)) {
          B200RT_CUDA(c, cudaMemcpyAsync(out, src, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        } else {
          tmp.resize(n);
          B200RT_CUDA(c, cudaMemcpyAsync(tmp.data(), src, (size_t) n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
          B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
          for (long long i = 0; i < n; i++) out[i] = tmp[i];
        }
      }
    row0 += rows[q];
  }
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  return B200RT_OK;
}

} // namespace

// ---- the two calls a device group hands its members with a slice of the caller's arrays (device_group.cu)
namespace b200rt {

int los_download_slice(b200rt_ctx *c, double *const dst[4], long long stride, long long offset) {
  if (!c) return B200RT_ERR_ARG;
  if (!c->los_done) return fail(c, B200RT_ERR_STATE, "no brightness result to download");
  cudaSetDevice(c->device);
  if (c->mult.defined)
    return is64(c) ? mult_los_download_impl<double>(c, dst, stride, offset) : mult_los_download_impl<float>(c, dst, stride, offset);
  return is64(c) ? los_download_impl<double>(c, dst, stride, offset) : los_download_impl<float>(c, dst, stride, offset);
}

int brightness_slice(b200rt_ctx *c, int n, const double *const src[9], int n_subsamples, double *const dst[4],
                     long long stride, long long offset) {
  if (!c) return B200RT_ERR_ARG;
  if (is64(c) && !c->mult.defined && n > 0 && c->have_grid && c->n_em >= 1) {
    // double singlet model: upload, kernels and download pipelined batch by batch
    for (int a = 0; a < 9; a++) if (!src[a]) return fail(c, B200RT_ERR_ARG, "null line-of-sight array");
    cudaSetDevice(c->device);
    HostLos io{n, src, dst, stride, offset};
    return brightness_impl<double>(c, n_subsamples, &io);
  }
  int rc = b200rt_los_upload(c, n, src[0], src[1], src[2], src[3], src[4], src[5], src[6], src[7], src[8]);
  if (rc) return rc;
  rc = b200rt_brightness_resident(c, n_subsamples);
  if (rc) return rc;
  return los_download_slice(c, dst, stride, offset);
}

} // namespace b200rt

// ====================================================================== C ABI
extern "C" {

int b200rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int b200rt_create(int device, int precision, b200rt_ctx **out) {
  if (!out || (precision != B200RT_F64 && precision != B200RT_F32)) return B200RT_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return B200RT_ERR_CUDA;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return B200RT_ERR_CUDA;
  b200rt_ctx *c = new (std::nothrow) b200rt_ctx;
  if (!c) return B200RT_ERR_NOMEM;
  c->device = device;
  c->precision = precision;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return B200RT_ERR_CUDA; }
  if (c->work_counter.ensure(4 * sizeof(int)) != cudaSuccess || c->step_counter.ensure(sizeof(unsigned long long)) != cudaSuccess) {
    delete c;
    return B200RT_ERR_CUDA;
  }
  *out = c;
  return B200RT_OK;
}

int b200rt_destroy(b200rt_ctx *c) {
  if (!c) return B200RT_OK;
  if (c->group) return group_destroy(c);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  DevBuf *bufs[] = {&c->grid_tables, &c->sun_rays, &c->list_dist, &c->list_ent, &c->list_len, &c->list_flag,
                    &c->work_counter, &c->step_counter, &c->los_in, &c->los_out, &c->los_order, &c->lu, &c->lu_dinv, &c->lu_flag,
                    &c->iph.dev, &c->iph.io, &c->vox_map, &c->sph_table};
  for (DevBuf *b : bufs) b->release();
  c->host_scratch.release();
  c->host_words.release();
  c->host_out.release();
  for (int e = 0; e < MAX_EMISSIONS; e++) {
    Emission &E = c->em[e];
    DevBuf *eb[] = {&E.tabs, &E.phi, &E.mrec, &E.K, &E.S0, &E.tau_sp, &E.tau_abs, &E.S, &E.S_real, &E.rec_pt, &E.rec_avg};
    for (DevBuf *b : eb) b->release();
  }
  {
    Multiplet &M = c->mult;
    DevBuf *mb[] = {&M.tabs, &M.rec_step, &M.rec_org, &M.rec_w0, &M.tsv, &M.tav, &M.K, &M.S0, &M.tau_sp, &M.tau_abs, &M.S,
                    &M.S_real, &M.rec_pt, &M.rec_avg};
    for (DevBuf *b : mb) b->release();
  }
  if (c->grid_view) ::operator delete(c->grid_view);
  for (cudaEvent_t ev : c->lu_events) cudaEventDestroy(ev);
  if (c->lu_graph) cudaGraphExecDestroy(c->lu_graph);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->out_stream) cudaStreamDestroy(c->out_stream);
  if (c->ev_rows) cudaEventDestroy(c->ev_rows);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream3) cudaStreamDestroy(c->stream3);
  cudaStreamDestroy(c->stream);
  delete c;
  return B200RT_OK;
}

const char *b200rt_last_error(const b200rt_ctx *c) { return c ? c->err.c_str() : "null context"; }

int b200rt_synchronize(b200rt_ctx *c) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_synchronize(c));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  return B200RT_OK;
}

int b200rt_make_grid_sph(int precision, int n_rb, int n_sb, int n_theta, int n_phi, const double *rb, int szamethod,
                         int raymethod, double *sb, double *pts_r, double *pts_s, double *ray_t, double *ray_p,
                         double *ray_domega) {
  if (n_rb < 2 || n_sb < 3 || n_theta < 2 || n_phi < 1 || !rb) return B200RT_ERR_ARG;
  if (precision == B200RT_F64) make_grid_sph<double>(n_rb, n_sb, n_theta, n_phi, rb, szamethod, raymethod, sb, pts_r, pts_s, ray_t, ray_p, ray_domega);
  else make_grid_sph<float>(n_rb, n_sb, n_theta, n_phi, rb, szamethod, raymethod, sb, pts_r, pts_s, ray_t, ray_p, ray_domega);
  return B200RT_OK;
}

int b200rt_set_grid_sph(b200rt_ctx *c, int n_rb, int n_sb, int n_rays, const double *rb, const double *sb,
                        const double *pts_r, const double *pts_s, const double *ray_t, const double *ray_p,
                        const double *ray_domega) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_grid_sph(c, n_rb, n_sb, n_rays, rb, sb, pts_r, pts_s, ray_t, ray_p, ray_domega));
  if (n_rb < 2 || n_sb < 3 || n_rays < 1 || !rb || !sb || !pts_r || !pts_s || !ray_t || !ray_p || !ray_domega)
    return fail(c, B200RT_ERR_ARG, "b200rt_set_grid_sph: bad argument");
  cudaSetDevice(c->device);
  HostGrid &h = c->hg;
  h.pp = false;
  h.n_rb = n_rb; h.n_sb = n_sb; h.n_vox = (n_rb - 1) * (n_sb - 1); h.n_rays = n_rays; h.cap = 2 * n_rb + n_sb;
  h.rb.assign(rb, rb + n_rb); h.sb.assign(sb, sb + n_sb);
  h.pts_r.assign(pts_r, pts_r + n_rb - 1); h.pts_s.assign(pts_s, pts_s + n_sb - 1);
  h.ray_t.assign(ray_t, ray_t + n_rays); h.ray_p.assign(ray_p, ray_p + n_rays);
  h.ray_domega.assign(ray_domega, ray_domega + n_rays);
  int rc = is64(c) ? upload_grid<double>(c) : upload_grid<float>(c);
  if (rc) return rc;
  c->have_grid = true;
  c->n_em = 0;
  for (int e = 0; e < 2; e++) { c->row_sink[e] = nullptr; c->row_sink_n_vox[e] = 0; }   // named for the old geometry
  for (int e = 0; e < MAX_EMISSIONS; e++) { c->em[e].defined = c->em[e].have_K = c->em[e].have_S = false; }
  c->mult.defined = c->mult.have_K = c->mult.have_S = false;
  return B200RT_OK;
}

int b200rt_make_grid_pp(int precision, int n_rb, int n_theta, const double *rb, double *pts_r, double *ray_t,
                        double *ray_domega) {
  if (n_rb < 2 || n_theta < 1 || !rb || !pts_r || !ray_t || !ray_domega) return B200RT_ERR_ARG;
  if (precision == B200RT_F64) make_grid_pp<double>(n_rb, n_theta, rb, pts_r, ray_t, ray_domega);
  else make_grid_pp<float>(n_rb, n_theta, rb, pts_r, ray_t, ray_domega);
  return B200RT_OK;
}

int b200rt_set_grid_pp(b200rt_ctx *c, int n_rb, int n_rays, const double *rb, const double *pts_r,
                       const double *ray_t, const double *ray_domega) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_grid_pp(c, n_rb, n_rays, rb, pts_r, ray_t, ray_domega));
  if (n_rb < 2 || n_rays < 1 || !rb || !pts_r || !ray_t || !ray_domega)
    return fail(c, B200RT_ERR_ARG, "b200rt_set_grid_pp: bad argument");
  cudaSetDevice(c->device);
  HostGrid &h = c->hg;
  // one SZA column [0, pi] whose voxel points sit on the +z axis (pt.xyz(0,0,pts_radii[i]) => pt.t = 0,
  // grid_plane_parallel.hpp:189-204); rays have phi = 0 (:219)
  h.pp = true;
  h.n_rb = n_rb; h.n_sb = 2; h.n_vox = n_rb - 1; h.n_rays = n_rays; h.cap = 2 * n_rb + 2;
  h.rb.assign(rb, rb + n_rb); h.sb.assign({0.0, M_PI});
  h.pts_r.assign(pts_r, pts_r + n_rb - 1); h.pts_s.assign(1, 0.0);
  h.ray_t.assign(ray_t, ray_t + n_rays); h.ray_p.assign(n_rays, 0.0);
  h.ray_domega.assign(ray_domega, ray_domega + n_rays);
  int rc = is64(c) ? upload_grid<double>(c) : upload_grid<float>(c);
  if (rc) return rc;
  c->have_grid = true;
  c->n_em = 0;
  for (int e = 0; e < 2; e++) { c->row_sink[e] = nullptr; c->row_sink_n_vox[e] = 0; }   // named for the old geometry
  for (int e = 0; e < MAX_EMISSIONS; e++) { c->em[e].defined = c->em[e].have_K = c->em[e].have_S = false; }
  c->mult.defined = c->mult.have_K = c->mult.have_S = false;
  return B200RT_OK;
}

int b200rt_set_singlet(b200rt_ctx *c, int e, int n_em, double branching, double T_ref, double sigma_ref, double g,
                       const double *T_ratio, const double *density, const double *dtau_species,
                       const double *dtau_absorber, const double *T_ratio_pt, const double *density_pt,
                       const double *dtau_species_pt, const double *dtau_absorber_pt) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_singlet(c, e, n_em, branching, T_ref, sigma_ref, g, T_ratio, density, dtau_species, dtau_absorber, T_ratio_pt, density_pt, dtau_species_pt, dtau_absorber_pt));
  if (!c->have_grid) return fail(c, B200RT_ERR_STATE, "set the grid before the emissions");
  if (n_em < 1 || n_em > MAX_EMISSIONS || e < 0 || e >= n_em) return fail(c, B200RT_ERR_ARG, "bad emission index");
  const double *arr[8] = {T_ratio, density, dtau_species, dtau_absorber, T_ratio_pt, density_pt, dtau_species_pt, dtau_absorber_pt};
  for (auto p : arr) if (!p) return fail(c, B200RT_ERR_ARG, "null emission table");
  cudaSetDevice(c->device);
  c->n_em = n_em;
  c->mult.defined = false;
  Emission &E = c->em[e];
  E.branching = branching; E.T_ref = T_ref; E.sigma_ref = sigma_ref; E.g_factor = g;
  int rc = is64(c) ? set_singlet_impl<double>(c, e, arr) : set_singlet_impl<float>(c, e, arr);
  if (rc) return rc;
  E.defined = true; E.have_K = false; E.have_S = false; E.residual = -1;
  return B200RT_OK;
}

int b200rt_multiplet_desc_init(int kind, int precision, b200rt_multiplet_desc *d) {
  if (!d || (precision != B200RT_F64 && precision != B200RT_F32)) return B200RT_ERR_ARG;
  return precision == B200RT_F64 ? multiplet_desc_init<double>(kind, d) : multiplet_desc_init<float>(kind, d);
}

int b200rt_set_multiplet(b200rt_ctx *c, const b200rt_multiplet_desc *d, const double *species_density,
                         const double *species_density_pt, const double *species_T, const double *species_T_pt,
                         const double *absorber_density, const double *absorber_density_pt) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_multiplet(c, d, species_density, species_density_pt, species_T, species_T_pt, absorber_density, absorber_density_pt));
  if (!c->have_grid) return fail(c, B200RT_ERR_STATE, "set the grid before the emissions");
  const double *arr[6] = {species_density, species_density_pt, species_T, species_T_pt, absorber_density, absorber_density_pt};
  for (auto p : arr) if (!p) return fail(c, B200RT_ERR_ARG, "null emission table");
  if (!d || mult_check_desc(*d))
    return fail(c, B200RT_ERR_ARG, "multiplet descriptor does not match the line / level tables of its kind");
  cudaSetDevice(c->device);
  c->n_em = 0;
  for (int e = 0; e < 2; e++) { c->row_sink[e] = nullptr; c->row_sink_n_vox[e] = 0; }
  for (int e = 0; e < MAX_EMISSIONS; e++) { c->em[e].defined = c->em[e].have_K = c->em[e].have_S = false; }
  Multiplet &M = c->mult;
  M.d = *d;
  M.defined = false;
  int rc = is64(c) ? set_multiplet_impl<double>(c, arr) : set_multiplet_impl<float>(c, arr);
  if (rc) return rc;
  M.defined = true; M.have_K = false; M.have_S = false; M.residual = -1; M.rec_dirty = true;
  return B200RT_OK;
}

int b200rt_set_g_factor(b200rt_ctx *c, int e, double g) {
  if (!c || e < 0 || e >= MAX_EMISSIONS) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_g_factor(c, e, g));
  c->em[e].g_factor = g;
  return B200RT_OK;
}

int b200rt_influence(b200rt_ctx *c, int v_begin, int v_end) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_influence(c, 1, &v_begin, &v_end));
  if (!c->have_grid || (c->n_em < 1 && !c->mult.defined)) return fail(c, B200RT_ERR_STATE, "grid / emissions not set");
  if (v_begin < 0 || v_end > c->hg.n_vox || v_begin > v_end) return fail(c, B200RT_ERR_ARG, "bad voxel range");
  cudaSetDevice(c->device);
  if (c->mult.defined) return is64(c) ? mult_influence_impl<double>(c, v_begin, v_end) : mult_influence_impl<float>(c, v_begin, v_end);
  const std::vector<std::pair<int, int>> one = {{v_begin, v_end}};
  return is64(c) ? influence_impl<double>(c, one) : influence_impl<float>(c, one);
}

int b200rt_influence_ranges(b200rt_ctx *c, int n_ranges, const int *v_begin, const int *v_end) {
  if (!c || n_ranges < 0 || (n_ranges > 0 && (!v_begin || !v_end))) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_influence(c, n_ranges, v_begin, v_end));
  if (!c->have_grid || c->n_em < 1) return fail(c, B200RT_ERR_STATE, "grid / singlet emissions not set");
  if (c->mult.defined) return fail(c, B200RT_ERR_STATE, "b200rt_influence_ranges: singlet emissions only");
  std::vector<std::pair<int, int>> r;
  int last_end = 0;
  for (int i = 0; i < n_ranges; i++) {
    if (v_begin[i] < last_end || v_end[i] > c->hg.n_vox || v_begin[i] > v_end[i])
      return fail(c, B200RT_ERR_ARG, "voxel ranges must be ascending, disjoint and inside the grid");
    last_end = v_end[i];
    r.emplace_back(v_begin[i], v_end[i]);
  }
  cudaSetDevice(c->device);
  return is64(c) ? influence_impl<double>(c, r) : influence_impl<float>(c, r);
}

int b200rt_solve(b200rt_ctx *c) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_solve(c));
  cudaSetDevice(c->device);
  if (c->mult.defined) return mult_solve_impl(c, true);
  return solve_impl(c, true);
}

int b200rt_generate_S(b200rt_ctx *c) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_generate_S(c));
  int rc = b200rt_influence(c, 0, c->hg.n_vox);
  if (rc) return rc;
  if (c->mult.defined) return mult_solve_impl(c, false);
  return solve_impl(c, false);
}

int b200rt_last_step_count(b200rt_ctx *c, long long *n) {
  if (!c || !n) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_counts(c, 0, n));
  *n = c->last_steps;
  return B200RT_OK;
}

int b200rt_last_substep_count(b200rt_ctx *c, long long *n) {
  if (!c || !n) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_counts(c, 1, n));
  *n = c->last_substeps;
  return B200RT_OK;
}

int b200rt_get_solution(b200rt_ctx *c, int e, double *S, double *S0, double *tsp, double *tab) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_get_solution(group_primary(c), e, S, S0, tsp, tab)));
  if (c->mult.defined) {
    if (e != 0) return B200RT_ERR_ARG;
    cudaSetDevice(c->device);
    Multiplet &M = c->mult;
    const size_t ne = (size_t) c->hg.n_vox * M.d.n_upper * sizeof(double), nl = (size_t) c->hg.n_vox * M.d.n_lines * sizeof(double);
    if (S) {
      if (!M.have_S) return fail(c, B200RT_ERR_STATE, "source function not solved");
      B200RT_CUDA(c, cudaMemcpyAsync(S, M.S.p, ne, cudaMemcpyDeviceToHost, c->stream));
    }
    if ((S0 || tsp || tab) && !M.have_K) return fail(c, B200RT_ERR_STATE, "influence pass not run");
    if (S0) B200RT_CUDA(c, cudaMemcpyAsync(S0, M.S0.p, ne, cudaMemcpyDeviceToHost, c->stream));
    if (tsp) B200RT_CUDA(c, cudaMemcpyAsync(tsp, M.tau_sp.p, nl, cudaMemcpyDeviceToHost, c->stream));
    if (tab) B200RT_CUDA(c, cudaMemcpyAsync(tab, M.tau_abs.p, nl, cudaMemcpyDeviceToHost, c->stream));
    B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
    return B200RT_OK;
  }
  if (e < 0 || e >= c->n_em) return B200RT_ERR_ARG;
  cudaSetDevice(c->device);
  Emission &E = c->em[e];
  const size_t nb = (size_t) c->hg.n_vox * sizeof(double);
  if (S) {
    if (!E.have_S) return fail(c, B200RT_ERR_STATE, "source function not solved");
    B200RT_CUDA(c, cudaMemcpyAsync(S, E.S.p, nb, cudaMemcpyDeviceToHost, c->stream));
  }
  if ((S0 || tsp || tab) && !E.have_K) return fail(c, B200RT_ERR_STATE, "influence pass not run");
  if (S0) B200RT_CUDA(c, cudaMemcpyAsync(S0, E.S0.p, nb, cudaMemcpyDeviceToHost, c->stream));
  if (tsp) B200RT_CUDA(c, cudaMemcpyAsync(tsp, E.tau_sp.p, nb, cudaMemcpyDeviceToHost, c->stream));
  if (tab) B200RT_CUDA(c, cudaMemcpyAsync(tab, E.tau_abs.p, nb, cudaMemcpyDeviceToHost, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  return B200RT_OK;
}

int b200rt_get_influence(b200rt_ctx *c, int e, int layout, double *K) {
  if (!c || !K) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_get_influence(group_primary(c), e, layout, K)));
  const bool mm = c->mult.defined;
  if (mm ? e != 0 : (e < 0 || e >= c->n_em)) return B200RT_ERR_ARG;
  cudaSetDevice(c->device);
  if (!(mm ? c->mult.have_K : c->em[e].have_K)) return fail(c, B200RT_ERR_STATE, "influence matrix not built");
  const size_t n = mm ? (size_t) c->hg.n_vox * c->mult.d.n_upper : (size_t) c->hg.n_vox;
  const void *Kdev = mm ? c->mult.K.p : c->em[e].K.p;
  if (layout == B200RT_ROW_MAJOR) {
    B200RT_CUDA(c, cudaMemcpyAsync(K, Kdev, n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  } else {
    std::vector<double> tmp(n * n);
    B200RT_CUDA(c, cudaMemcpyAsync(tmp.data(), Kdev, n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) for (size_t j = 0; j < n; j++) K[j * n + i] = tmp[i * n + j];
  }
  return B200RT_OK;
}

int b200rt_set_sourcefn(b200rt_ctx *c, int e, const double *S) {
  if (!c || !S) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_set_sourcefn(c, e, S));
  if (c->mult.defined) {
    if (e != 0) return B200RT_ERR_ARG;
    cudaSetDevice(c->device);
    Multiplet &M = c->mult;
    const int ne = c->hg.n_vox * M.d.n_upper;
    B200RT_CUDA(c, cudaMemcpyAsync(M.S.p, S, ne * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (is64(c)) B200RT_CUDA(c, launch_convert<double>(M.S.as<double>(), M.S_real.as<double>(), ne, c->stream));
    else B200RT_CUDA(c, launch_convert<float>(M.S.as<double>(), M.S_real.as<float>(), ne, c->stream));
    B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
    M.have_S = true;
    M.rec_dirty = true;
    return B200RT_OK;
  }
  if (e < 0 || e >= c->n_em) return B200RT_ERR_ARG;
  cudaSetDevice(c->device);
  Emission &E = c->em[e];
  if (!E.defined) return fail(c, B200RT_ERR_STATE, "emission not defined");
  const int n = c->hg.n_vox;
  B200RT_CUDA(c, cudaMemcpyAsync(E.S.p, S, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (is64(c)) B200RT_CUDA(c, launch_convert<double>(E.S.as<double>(), E.S_real.as<double>(), n, c->stream));
  else B200RT_CUDA(c, launch_convert<float>(E.S.as<double>(), E.S_real.as<float>(), n, c->stream));
  B200RT_CUDA(c, cudaStreamSynchronize(c->stream));
  E.have_S = true;
  E.rec_dirty = true;
  return B200RT_OK;
}

int b200rt_last_residual(b200rt_ctx *c, int e, double *r) {
  if (!c || !r) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_last_residual(group_primary(c), e, r)));
  if (c->mult.defined) { *r = c->mult.residual; return e == 0 ? B200RT_OK : B200RT_ERR_ARG; }
  if (e < 0 || e >= c->n_em) return B200RT_ERR_ARG;
  *r = c->em[e].residual;
  return B200RT_OK;
}

int b200rt_influence_dev(b200rt_ctx *c, int e, void **K, void **S0, void **tsp, void **tab) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_influence_dev(group_primary(c), e, K, S0, tsp, tab)));
  if (c->mult.defined) {
    if (e != 0) return B200RT_ERR_ARG;
    Multiplet &M = c->mult;
    if (K) *K = M.K.p;
    if (S0) *S0 = M.S0.p;
    if (tsp) *tsp = M.tau_sp.p;
    if (tab) *tab = M.tau_abs.p;
    M.have_K = true;
    return B200RT_OK;
  }
  if (e < 0 || e >= c->n_em) return B200RT_ERR_ARG;
  Emission &E = c->em[e];
  if (!E.defined) return fail(c, B200RT_ERR_STATE, "emission not defined");
  if (K) *K = E.K.p;
  if (S0) *S0 = E.S0.p;
  if (tsp) *tsp = E.tau_sp.p;
  if (tab) *tab = E.tau_abs.p;
  E.have_K = true;   // the caller may fill rows from peers
  return B200RT_OK;
}

// ---- peer-memory row exchange (one process per GPU): the solving rank exports its K, the others open it and
// name it as the sink of their row batches
int b200rt_ipc_export_influence(b200rt_ctx *c, int e, void *handle64) {
  if (c && c->group) return group_forward(c, b200rt_ipc_export_influence(group_primary(c), e, handle64));
  if (!c || !handle64 || e < 0 || e >= c->n_em || c->mult.defined) return B200RT_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  cudaSetDevice(c->device);
  cudaIpcMemHandle_t h;
  B200RT_CUDA(c, cudaIpcGetMemHandle(&h, c->em[e].K.p));
  std::memcpy(handle64, &h, sizeof h);
  c->em[e].have_K = true;   // peers fill rows
  return B200RT_OK;
}
int b200rt_ipc_open(b200rt_ctx *c, const void *handle64, void **peer_ptr) {
  if (!c || !handle64 || !peer_ptr) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_ipc_open(group_primary(c), handle64, peer_ptr)));
  cudaSetDevice(c->device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof h);
  B200RT_CUDA(c, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return B200RT_OK;
}
int b200rt_ipc_close(b200rt_ctx *c, void *peer_ptr) {
  if (!c || !peer_ptr) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_ipc_close(group_primary(c), peer_ptr)));
  cudaSetDevice(c->device);
  for (auto &sk : c->row_sink) if (sk == peer_ptr) sk = nullptr;
  B200RT_CUDA(c, cudaIpcCloseMemHandle(peer_ptr));
  return B200RT_OK;
}
int b200rt_set_row_sink(b200rt_ctx *c, int e, void *peer_K_dev) {
  if (!c || e < 0 || e >= 2) return B200RT_ERR_ARG;
  if (c->group) return fail(c, B200RT_ERR_STATE, "b200rt_set_row_sink: a device group wires its own row sinks");
  if (peer_K_dev) {
    if (c->mult.defined) return fail(c, B200RT_ERR_STATE, "b200rt_set_row_sink: singlet emissions only");
    if (!c->have_grid || e >= c->n_em) return fail(c, B200RT_ERR_STATE, "b200rt_set_row_sink: set the grid and the emissions first");
  }
  c->row_sink[e] = peer_K_dev;
  c->row_sink_n_vox[e] = peer_K_dev ? c->hg.n_vox : 0;
  return B200RT_OK;
}

int b200rt_sourcefn_dev(b200rt_ctx *c, int e, void **S) {
  if (c && c->group) return group_forward(c, b200rt_sourcefn_dev(group_primary(c), e, S));
  if (c && S && c->mult.defined && e == 0) { *S = c->mult.S.p; return B200RT_OK; }
  if (!c || e < 0 || e >= c->n_em || !S) return B200RT_ERR_ARG;
  *S = c->em[e].S.p;
  return B200RT_OK;
}

int b200rt_los_from_MSO(int precision, int n, const double *loc, const double *dir, double *x, double *y, double *z,
                        double *r, double *t, double *lx, double *ly, double *lz, double *cost) {
  if (n < 0 || !loc || !dir || !x || !y || !z || !r || !t || !lx || !ly || !lz || !cost) return B200RT_ERR_ARG;
  if (precision == B200RT_F64) los_from_MSO<double>(n, loc, dir, x, y, z, r, t, lx, ly, lz, cost);
  else los_from_MSO<float>(n, loc, dir, x, y, z, r, t, lx, ly, lz, cost);
  return B200RT_OK;
}

int b200rt_los_upload(b200rt_ctx *c, int n, const double *x, const double *y, const double *z, const double *r,
                      const double *t, const double *lx, const double *ly, const double *lz, const double *cost) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_los_upload(c, n, x, y, z, r, t, lx, ly, lz, cost));
  if (n <= 0) return fail(c, B200RT_ERR_ARG, "there must be at least one observation to simulate");
  const double *src[9] = {x, y, z, r, t, lx, ly, lz, cost};
  for (auto p : src) if (!p) return fail(c, B200RT_ERR_ARG, "null line-of-sight array");
  cudaSetDevice(c->device);
  return is64(c) ? los_upload_impl<double>(c, n, src) : los_upload_impl<float>(c, n, src);
}

int b200rt_brightness_resident(b200rt_ctx *c, int n_subsamples) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_brightness_resident(c, n_subsamples));
  if (!c->have_grid || (c->n_em < 1 && !c->mult.defined)) return fail(c, B200RT_ERR_STATE, "grid / emissions not set");
  cudaSetDevice(c->device);
  if (c->mult.defined) return is64(c) ? mult_brightness_impl<double>(c, n_subsamples) : mult_brightness_impl<float>(c, n_subsamples);
  return is64(c) ? brightness_impl<double>(c, n_subsamples) : brightness_impl<float>(c, n_subsamples);
}

int b200rt_los_download(b200rt_ctx *c, double *B, double *tsp, double *tab, double *col) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_los_download(c, B, tsp, tab, col));
  double *dst[4] = {B, tsp, tab, col};
  return los_download_slice(c, dst, c->n_los, 0);
}

int b200rt_brightness(b200rt_ctx *c, int n, const double *x, const double *y, const double *z, const double *r,
                      const double *t, const double *lx, const double *ly, const double *lz, const double *cost,
                      int n_subsamples, double *B, double *tsp, double *tab, double *col) {
  if (!c) return B200RT_ERR_ARG;
  const double *src[9] = {x, y, z, r, t, lx, ly, lz, cost};
  double *dst[4] = {B, tsp, tab, col};
  GROUP_DISPATCH(c, group_brightness(c, n, src, n_subsamples, dst));
  return brightness_slice(c, n, src, n_subsamples, dst, n, 0);
}

int b200rt_traverse_voxel_rays(b200rt_ctx *c, int v_begin, int v_end, long long capacity, int *len, int *eb,
                               int *entering, double *distance, long long *n_entries) {
  if (!c || !len || !eb || !entering || !distance) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_traverse_voxel_rays(group_primary(c), v_begin, v_end, capacity, len, eb, entering, distance, n_entries)));
  if (!c->have_grid) return fail(c, B200RT_ERR_STATE, "grid not set");
  if (v_begin < 0 || v_end > c->hg.n_vox || v_begin > v_end) return fail(c, B200RT_ERR_ARG, "bad voxel range");
  cudaSetDevice(c->device);
  return is64(c) ? traverse_voxel_rays_impl<double>(c, v_begin, v_end, capacity, len, eb, entering, distance, n_entries)
                 : traverse_voxel_rays_impl<float>(c, v_begin, v_end, capacity, len, eb, entering, distance, n_entries);
}

int b200rt_traverse_los(b200rt_ctx *c, long long capacity, int *len, int *eb, int *entering, double *distance,
                        long long *n_entries) {
  if (!c || !len || !eb || !entering || !distance) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_traverse_los(c, capacity, len, eb, entering, distance, n_entries));
  if (!c->have_grid) return fail(c, B200RT_ERR_STATE, "grid not set");
  cudaSetDevice(c->device);
  return is64(c) ? traverse_los_impl<double>(c, capacity, len, eb, entering, distance, n_entries)
                 : traverse_los_impl<float>(c, capacity, len, eb, entering, distance, n_entries);
}

int b200rt_last_kernel_ms(b200rt_ctx *c, int phase, float *ms, int *n_launches) {
  if (!c || phase < 0 || phase >= PH_COUNT) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_kernel_ms(c, phase, ms, n_launches));
  if (ms) *ms = c->phase_ms[phase];
  if (n_launches) *n_launches = c->phase_launches[phase];
  return B200RT_OK;
}

int b200rt_iph_load_table(b200rt_ctx *c, const char *fname) {
  if (!c || !fname) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_iph_load_table(c, fname));
  cudaSetDevice(c->device);
  return iph_load_table(c, fname);
}

int b200rt_iph_set_table(b200rt_ctx *c, int kmax, int lmax, int ninf, float temp, const float *alt_au, const float *ang,
                         const float *dans, const float *sot, const float *so, const float *sn, const float *dinf_cm3) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_iph_set_table(c, kmax, lmax, ninf, temp, alt_au, ang, dans, sot, so, sn, dinf_cm3));
  cudaSetDevice(c->device);
  return iph_set_table(c, kmax, lmax, ninf, temp, alt_au, ang, dans, sot, so, sn, dinf_cm3);
}

int b200rt_iph_background(b200rt_ctx *c, float fs, float xpos, float ypos, float zpos, int n_los, const float *u,
                          const float *v, const float *w, float *fln, int *n_steps) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_iph_background(c, fs, xpos, ypos, zpos, n_los, u, v, w, fln, n_steps));
  cudaSetDevice(c->device);
  return iph_background(c, fs, xpos, ypos, zpos, n_los, u, v, w, fln, n_steps);
}

int b200rt_iph_model(b200rt_ctx *c, double g_lya, const double *marspos, int n_los, const double *ra, const double *dec,
                     double *iph_kR) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_iph_model(c, g_lya, marspos, n_los, ra, dec, iph_kR));
  cudaSetDevice(c->device);
  return iph_model(c, g_lya, marspos, n_los, ra, dec, iph_kR);
}

int b200rt_iph_extinction(int n, const double *iph, const double *tau_abs, double *out) {
  if (n < 0 || !iph || !tau_abs || !out) return B200RT_ERR_ARG;
  for (int i = 0; i < n; i++) out[i] = (tau_abs[i] != -1) ? iph[i] * std::exp(-tau_abs[i]) : 0.0;
  return B200RT_OK;
}

int b200rt_measure_fp64_peaks(b200rt_ctx *c, double *dfma, double *dmma) {
  if (!c) return B200RT_ERR_ARG;
  GROUP_DISPATCH(c, group_forward(c, b200rt_measure_fp64_peaks(group_primary(c), dfma, dmma)));
  cudaSetDevice(c->device);
  return measure_fp64_peaks(c, dfma, dmma);
}

} // extern "C"
