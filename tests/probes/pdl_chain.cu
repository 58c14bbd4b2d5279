// Micro-probe: cost per dependent small kernel inside a CUDA graph, with and without programmatic
// dependent launch (griddepcontrol).  nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_chain pdl_chain.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

template <bool PDL>
__global__ void k_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, long n) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = a[i] + b[i];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <bool PDL>
float run(long n, int chain, int reps) {
  float *a, *b, *c;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, n * 4));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
  for (int i = 0; i < chain; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((n + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    float* src = (i & 1) ? c : a; float* dst = (i & 1) ? a : c;
    CK(cudaLaunchKernelEx(&cfg, k_add<PDL>, (const float*)src, (const float*)b, dst, n));
  }
  CK(cudaStreamEndCapture(s, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(ge, s));
  CK(cudaEventRecord(e1, s));
  CK(cudaStreamSynchronize(s));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(a); cudaFree(b); cudaFree(c);
  return ms * 1000.f / reps / chain;
}

int main() {
  for (long n : {1024L, 655360L, 4194304L}) {
    float t0 = run<false>(n, 200, 20), t1 = run<true>(n, 200, 20);
    printf("n=%ld: plain %.2f us/kernel, pdl %.2f us/kernel\n", n, t0, t1);
  }
  return 0;
}
