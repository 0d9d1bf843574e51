#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *ctr, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0) { int v = ++(*ctr); if (v >= 10) cudaGraphSetConditional(h, 0); }
}
int main() {
  int *ctr; cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  printf("%d\n", (int) cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  printf("%d\n", (int) cudaGraphAddNode(&node, g, nullptr, 0, &p));
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  printf("%d\n", (int) cudaStreamBeginCaptureToGraph(s, bg, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  body<<<1, 32, 0, s>>>(ctr, h);
  cudaGraph_t tmp; printf("%d\n", (int) cudaStreamEndCapture(s, &tmp));
  cudaGraphExec_t ex; printf("%d\n", (int) cudaGraphInstantiate(&ex, g, 0));
  printf("%d\n", (int) cudaGraphLaunch(ex, s)); cudaStreamSynchronize(s);
  int v; cudaMemcpy(&v, ctr, 4, cudaMemcpyDeviceToHost); printf("ctr %d\n", v);
}
