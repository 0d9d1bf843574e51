// api_traverse.cu -- the traversal parity surface: boundary lists of voxel-origin rays / lines of sight back to the host
// (one of the translation units behind the C ABI; see api_internal.hpp)
#include "api_internal.hpp"

namespace b200rt {
namespace api {
namespace {

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


}  // namespace

int traverse_voxel_rays(b200rt_ctx *c, int v_begin, int v_end, long long capacity, int *len, int *exits_bottom,
                        int *entering, double *distance, long long *n_entries) {
  return is64(c) ? traverse_voxel_rays_impl<double>(c, v_begin, v_end, capacity, len, exits_bottom, entering, distance, n_entries)
                 : traverse_voxel_rays_impl<float>(c, v_begin, v_end, capacity, len, exits_bottom, entering, distance, n_entries);
}
int traverse_los(b200rt_ctx *c, long long capacity, int *len, int *exits_bottom, int *entering, double *distance,
                 long long *n_entries) {
  return is64(c) ? traverse_los_impl<double>(c, capacity, len, exits_bottom, entering, distance, n_entries)
                 : traverse_los_impl<float>(c, capacity, len, exits_bottom, entering, distance, n_entries);
}

}  // namespace api
}  // namespace b200rt
