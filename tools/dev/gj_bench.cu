// Development aid: times gj128_kernel alone on a diagonally dominant block.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3d_planetary_rt_model_b200/csrc -Iinclude tools/dev/gj_bench.cu -o tools/dev/bin/gj_bench
#include "../../3d_planetary_rt_model_b200/csrc/solve.cu"
#include <cstdio>
#include <cstdlib>
namespace b200rt {
int run() {
  const int np = 128;
  std::vector<double> h((size_t) np * np);
  for (int i = 0; i < np; i++) {
    double s = 0;
    for (int j = 0; j < np; j++) if (j != i) { h[(size_t) i * np + j] = -((i * 131 + j * 71) % 97) / 97.0 / np; s += -h[(size_t) i * np + j]; }
    h[(size_t) i * np + i] = 1.0 + 0.1 * s;
  }
  double *A, *dinv;
  cudaMalloc(&A, h.size() * 8); cudaMalloc(&dinv, h.size() * 8);
  cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; i++) gj128_kernel<<<1, GJ_THREADS>>>(A, np, 0, dinv);
  cudaEventRecord(e0);
  const int reps = 50;
  for (int i = 0; i < reps; i++) gj128_kernel<<<1, GJ_THREADS>>>(A, np, 0, dinv);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<double> inv(h.size());
  cudaMemcpy(inv.data(), dinv, h.size() * 8, cudaMemcpyDeviceToHost);
  double err = 0;
  for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) {
    double s = 0; for (int k = 0; k < np; k++) s += h[(size_t) i * np + k] * inv[(size_t) k * np + j];
    err = std::max(err, std::fabs(s - (i == j)));
  }
  if (getenv("GJ_MANY")) {   // 148 CTAs on 148 diagonal blocks: gives the profiler's sampler something to see
    const int nb = 148, npb = np * nb;
    double *Ab, *db;
    cudaMalloc(&Ab, (size_t) npb * npb * 8); cudaMalloc(&db, (size_t) nb * np * np * 8);
    cudaMemset(Ab, 0, (size_t) npb * npb * 8);
    for (int k = 0; k < nb; k++) cudaMemcpy2D(Ab + ((size_t) k * np) * npb + (size_t) k * np, (size_t) npb * 8, A, (size_t) np * 8, (size_t) np * 8, np, cudaMemcpyDeviceToDevice);
    for (int i = 0; i < 3; i++) gj128_kernel<<<nb, GJ_THREADS>>>(Ab, npb, 0, db);
    cudaDeviceSynchronize();
    printf("many: %s\n", cudaGetErrorString(cudaGetLastError()));
  }
  auto run_cluster = [&](int ncta) {
    auto launch = [&]() {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ncta); cfg.blockDim = dim3(GJC_THREADS);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ncta; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      const double *Ac = A; int npc = np, kb = 0; double *dc = dinv;
      if (ncta == 4) cudaLaunchKernelEx(&cfg, gj128_cluster_kernel<4>, Ac, npc, kb, dc);
      else cudaLaunchKernelEx(&cfg, gj128_cluster_kernel<8>, Ac, npc, kb, dc);
    };
    for (int i = 0; i < 5; i++) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float msc; cudaEventElapsedTime(&msc, e0, e1);
    std::vector<double> invc(h.size());
    cudaMemcpy(invc.data(), dinv, h.size() * 8, cudaMemcpyDeviceToHost);
    double errc = 0;
    for (int i = 0; i < np; i++) for (int j = 0; j < np; j++) {
      double s2 = 0; for (int k = 0; k < np; k++) s2 += h[(size_t) i * np + k] * invc[(size_t) k * np + j];
      errc = std::max(errc, std::fabs(s2 - (i == j)));
    }
    printf("cluster of %d: %.2f us per inverse (incl. launch), |A inv - I| max %.2e  %s\n", ncta, msc * 1e3 / reps, errc, cudaGetErrorString(cudaGetLastError()));
    cudaMemset(dinv, 0, h.size() * 8);
  };
  run_cluster(4);
  run_cluster(8);
  gj128_kernel<<<1, GJ_THREADS>>>(A, np, 0, dinv);
  cudaMemcpy(inv.data(), dinv, h.size() * 8, cudaMemcpyDeviceToHost);
  printf("%.2f us per inverse (incl. launch), |A inv - I| max %.2e  %s\n", ms * 1e3 / reps, err, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
}
int main() { return b200rt::run(); }
