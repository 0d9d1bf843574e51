// Development aid: latency / issue-rate probes for the FP64 paths of one SM (DMMA.8x8x4, DFMA, LDS, bar.sync).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// mode 0: dependent DMMA chain; 1: 16 independent DMMA accumulators; 2: dependent DFMA chain; 3: LDS pointer chase; 4: barriers
template <int MODE> __global__ void probe(double *out, long long *cyc, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double) ((i * 8 + 8) % 8192);
  __syncthreads();
  double c[16][2];
  for (int i = 0; i < 16; i++) { c[i][0] = threadIdx.x; c[i][1] = 1; }
  double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
  long long t0 = clock64();
  if (MODE == 0) for (int i = 0; i < iters; i++) dmma(c[0][0], c[0][1], a, b);
  if (MODE == 1) for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) dmma(c[k][0], c[k][1], a, b);
  }
  if (MODE == 2) for (int i = 0; i < iters; i++) c[0][0] = fma(c[0][0], b, a);
  if (MODE == 3) { int idx = threadIdx.x & 1023; for (int i = 0; i < iters; i++) idx = ((int) sm[idx]) >> 3; c[0][0] = idx; }
  if (MODE == 4) for (int i = 0; i < iters; i++) __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  const char *names[5] = {"dependent DMMA chain", "16 independent DMMAs", "dependent DFMA chain", "LDS chase", "bar.sync"};
  for (int threads : {32, 128, 256, 512}) {
    for (int mode = 0; mode < 5; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) probe<0><<<1, threads>>>(out, cyc, iters);
        if (mode == 1) probe<1><<<1, threads>>>(out, cyc, iters);
        if (mode == 2) probe<2><<<1, threads>>>(out, cyc, iters);
        if (mode == 3) probe<3><<<1, threads>>>(out, cyc, iters);
        if (mode == 4) probe<4><<<1, threads>>>(out, cyc, iters);
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double per = (double) h / iters / (mode == 1 ? 16 : 1);
      printf("threads %3d  %-22s %8.1f cycles per op (warp 0)\n", threads, names[mode], per);
    }
  }
  return 0;
}
