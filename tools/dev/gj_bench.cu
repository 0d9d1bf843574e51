// Development aid: times gj128_kernel alone on a diagonally dominant block.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I3d_planetary_rt_model_b200/csrc -Iinclude tools/dev/gj_bench.cu -o tools/dev/bin/gj_bench
#include "../../3d_planetary_rt_model_b200/csrc/solve.cu"
#include <cstdio>
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
  printf("%.2f us per inverse (incl. launch), |A inv - I| max %.2e  %s\n", ms * 1e3 / reps, err, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
}
int main() { return b200rt::run(); }
