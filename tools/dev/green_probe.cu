// Development aid: which runtime-API operations work on a green-context stream, and with which context current.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
template <class F> bool fn(const char *name, F &f) {
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return false;
  f = reinterpret_cast<F>(p); return true;
}
__global__ void k(int *x) { atomicAdd(x, 1); }
#define SHOW(what, call) do { cudaError_t e = (call); printf("%-58s %s\n", what, cudaGetErrorString(e)); cudaGetLastError(); } while (0)
int main() {
  cudaSetDevice(0); cudaFree(0);
  CUresult (*dget)(CUdevice *, int); CUresult (*getres)(CUdevice, CUdevResource *, CUdevResourceType);
  CUresult (*split)(CUdevResource *, unsigned *, const CUdevResource *, CUdevResource *, unsigned, unsigned);
  CUresult (*gdesc)(CUdevResourceDesc *, CUdevResource *, unsigned); CUresult (*gcreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned);
  CUresult (*gstream)(CUstream *, CUgreenCtx, unsigned, int); CUresult (*from)(CUcontext *, CUgreenCtx); CUresult (*setc)(CUcontext);
  CUresult (*getc)(CUcontext *);
  fn("cuDeviceGet", dget); fn("cuDeviceGetDevResource", getres); fn("cuDevSmResourceSplitByCount", split); fn("cuDevResourceGenerateDesc", gdesc);
  fn("cuGreenCtxCreate", gcreate); fn("cuGreenCtxStreamCreate", gstream); fn("cuCtxFromGreenCtx", from); fn("cuCtxSetCurrent", setc); fn("cuCtxGetCurrent", getc);
  CUdevice dev; dget(&dev, 0);
  CUdevResource sm; printf("getres %d\n", getres(dev, &sm, CU_DEV_RESOURCE_TYPE_SM)); printf("SMs %u\n", sm.sm.smCount);
  unsigned nb = 4; std::vector<CUdevResource> g(4); CUdevResource rest;
  printf("split %d -> %u groups of %u, rest %u\n", split(g.data(), &nb, &sm, &rest, 0, 32), nb, g[0].sm.smCount, rest.sm.smCount);
  CUdevResourceDesc desc; gdesc(&desc, &g[1], 1);
  CUgreenCtx gc; printf("gcreate %d\n", gcreate(&gc, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM));
  CUstream gs; printf("gstream %d\n", gstream(&gs, gc, CU_STREAM_NON_BLOCKING, 0));
  CUcontext primary; getc(&primary);
  CUcontext gctx; printf("from %d\n", from(&gctx, gc));
  int *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  int h = 0;
  printf("--- primary context current\n");
  k<<<1, 32, 0, gs>>>(d); SHOW("kernel launch on green stream", cudaGetLastError());
  SHOW("cudaMemcpyAsync D2H on green stream", cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, gs));
  SHOW("cudaMemsetAsync on green stream", cudaMemsetAsync(d, 0, 4, gs));
  SHOW("cudaStreamSynchronize(green stream)", cudaStreamSynchronize(gs));
  cudaEvent_t ep; cudaEventCreate(&ep);
  SHOW("cudaEventRecord(primary-created event, green stream)", cudaEventRecord(ep, gs));
  cudaStream_t ps; cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking);
  SHOW("cudaEventRecord(primary event, primary stream)", cudaEventRecord(ep, ps));
  SHOW("cudaStreamWaitEvent(green stream, primary event)", cudaStreamWaitEvent(gs, ep, 0));
  setc(gctx);
  cudaEvent_t eg, eg2; SHOW("[green current] cudaEventCreate", cudaEventCreate(&eg)); cudaEventCreate(&eg2);
  setc(primary);
  SHOW("cudaEventRecord(green-created event, green stream)", cudaEventRecord(eg, gs));
  k<<<1, 32, 0, gs>>>(d);
  SHOW("cudaEventRecord(green event 2, green stream)", cudaEventRecord(eg2, gs));
  SHOW("cudaEventSynchronize(green event 2)", cudaEventSynchronize(eg2));
  float ms = -1; SHOW("cudaEventElapsedTime(green events)", cudaEventElapsedTime(&ms, eg, eg2)); printf("   ms %f\n", ms);
  SHOW("cudaStreamWaitEvent(primary stream, green event)", cudaStreamWaitEvent(ps, eg2, 0));
  cudaGraph_t graph; cudaGraphExec_t ge;
  SHOW("cudaStreamBeginCapture(green stream)", cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed));
  k<<<1, 32, 0, gs>>>(d);
  SHOW("cudaStreamEndCapture", cudaStreamEndCapture(gs, &graph));
  SHOW("cudaGraphInstantiate", cudaGraphInstantiate(&ge, graph, 0));
  SHOW("cudaGraphLaunch(green stream)", cudaGraphLaunch(ge, gs));
  SHOW("cudaStreamSynchronize(green stream)", cudaStreamSynchronize(gs));
  printf("--- green context current\n");
  setc(gctx);
  k<<<1, 32, 0, gs>>>(d); SHOW("kernel launch on green stream", cudaGetLastError());
  SHOW("cudaMemcpyAsync D2H on green stream", cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, gs));
  SHOW("cudaStreamSynchronize(green stream)", cudaStreamSynchronize(gs));
  int *d2 = nullptr; SHOW("cudaMalloc", cudaMalloc(&d2, 1024));
  SHOW("cudaFuncSetAttribute", cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024));
  printf("counter %d\n", h);
  return 0;
}
