// Micro-benchmark: cost of a kernel boundary in a chain of tiny dependent kernels, with and without programmatic dependent launch.
// nvcc -arch=sm_100a -O3 -o pdl_chain pdl_chain.cu && ./pdl_chain
#include <cstdio>
#include <cuda_runtime.h>
__global__ void tiny(float* x, int pdl) {
  if (pdl) { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); asm volatile("griddepcontrol.wait;" ::: "memory"); }
  if (threadIdx.x == 0 && blockIdx.x == 0) x[0] += 1.0f;
}
static void run(int grid, int block, size_t smem, int pdl, int n) {
  float* x; cudaMalloc(&x, 4); cudaMemset(x, 0, 4);
  cudaFuncSetAttribute(tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(a);
    for (int i = 0; i < n; ++i) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      cudaLaunchKernelEx(&cfg, tiny, x, pdl);
    }
    cudaEventRecord(b); cudaEventSynchronize(b);
  }
  float ms; cudaEventElapsedTime(&ms, a, b);
  printf("grid %4d block %4d smem %6zu pdl %d : %.2f us per kernel\n", grid, block, smem, pdl, ms * 1000.f / n);
  cudaFree(x);
}
int main() {
  for (int pdl = 0; pdl < 2; ++pdl) {
    run(1, 256, 0, pdl, 2000);
    run(296, 288, 100 * 1024, pdl, 2000);
    run(256, 256, 96 * 1024, pdl, 2000);
  }
  return 0;
}
