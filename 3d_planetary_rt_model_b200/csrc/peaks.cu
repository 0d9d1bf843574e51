// peaks.cu -- micro-benchmarks of the pipes that bound this hot path.
//
// MEASURED_PEAKS.json (driver-written) holds HBM copy bandwidth and dense bf16 only; the ray
// march and the brightness march are FP64-instruction bound and the solve runs on the FP64
// tensor pipe, so their roofline denominators are measured here, on the box, in the same run:
//   DFMA  : 8 independent fused multiply-add chains per thread
//   DMMA  : mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4), 8 independent accumulator pairs
//   MUFU  : ex2.approx.f32 throughput is not needed (the f64 exp is DFMA-based)
#include "common.hpp"

namespace b200rt {
namespace {

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678) out[0] = x0;   // keep the chains alive
}

__global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; j++) { c[j][0] = threadIdx.x + j; c[j][1] = j; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
  if (s == 12345.678) out[0] = s;
}

} // namespace

int measure_fp64_peaks(b200rt_ctx *c, double *dfma_tflops, double *dmma_tflops) {
  DevBuf sink;
  B200RT_CUDA(c, sink.ensure(64));
  cudaEvent_t e0, e1;
  B200RT_CUDA(c, cudaEventCreate(&e0));
  B200RT_CUDA(c, cudaEventCreate(&e1));
  const int blocks = NUM_SMS * 8, threads = 256, iters = 1 << 14;
  float best_fma = 1e30f, best_mma = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0, c->stream);
    dfma_kernel<<<blocks, threads, 0, c->stream>>>(sink.as<double>(), iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    B200RT_CUDA(c, cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_fma) best_fma = ms;
    cudaEventRecord(e0, c->stream);
    dmma_kernel<<<blocks, threads, 0, c->stream>>>(sink.as<double>(), iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    B200RT_CUDA(c, cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_mma) best_mma = ms;
  }
  const double fma_flop = 2.0 * 8 * (double) iters * blocks * threads;
  const double mma_flop = 2.0 * 256 * 8 * (double) iters * blocks * (threads / 32);   // 8x8x4 MACs per warp instruction
  if (dfma_tflops) *dfma_tflops = fma_flop / (best_fma * 1e-3) / 1e12;
  if (dmma_tflops) *dmma_tflops = mma_flop / (best_mma * 1e-3) / 1e12;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  sink.release();
  return B200RT_OK;
}

} // namespace b200rt
