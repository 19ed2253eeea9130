// Probe: device time between two events around an (almost) empty kernel, for a plain and a
// cooperative launch, with and without a large dynamic shared memory request.
#include <cstdio>
#include <cuda_runtime.h>
#include <algorithm>
#include <vector>
__global__ void k_empty(int* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) *p = 1; }
int main() {
  int* d; cudaMalloc(&d, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int coop = 0; coop < 2; ++coop)
    for (size_t smem : {size_t(0), size_t(156 * 1024)})
      for (int threads : {256, 1024}) {
        std::vector<float> t;
        for (int it = 0; it < 30; ++it) {
          void* args[] = {(void*)&d};
          cudaEventRecord(e0, s);
          if (coop) cudaLaunchCooperativeKernel((const void*)k_empty, dim3(148), dim3(threads), args, smem, s);
          else k_empty<<<148, threads, smem, s>>>(d);
          cudaEventRecord(e1, s);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          t.push_back(ms * 1e3f);
        }
        std::sort(t.begin(), t.end());
        printf("coop=%d smem=%zu threads=%d: min %.2f med %.2f us (%s)\n", coop, smem, threads, t[0], t[t.size() / 2], cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
