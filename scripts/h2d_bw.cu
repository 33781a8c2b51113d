// H2D bandwidth probe: default pinned vs write-combined pinned, 1 and 2 streams.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstring>
int main() {
  size_t n = (size_t)4 << 30;
  void *d, *h1, *h2;
  cudaMalloc(&d, n);
  cudaMallocHost(&h1, n);
  cudaHostAlloc(&h2, n, cudaHostAllocWriteCombined);
  memset(h1, 1, n); memset(h2, 1, n);
  cudaStream_t s[2]; cudaStreamCreate(&s[0]); cudaStreamCreate(&s[1]);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 4; mode++) {
    void* h = (mode & 1) ? h2 : h1; int ns = (mode & 2) ? 2 : 1;
    float best = 1e9;
    for (int it = 0; it < 4; it++) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, s[0]);
      if (ns == 1) cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s[0]);
      else {
        cudaStreamWaitEvent(s[1], e0, 0);
        cudaMemcpyAsync(d, h, n / 2, cudaMemcpyHostToDevice, s[0]);
        cudaMemcpyAsync((char*)d + n / 2, (char*)h + n / 2, n / 2, cudaMemcpyHostToDevice, s[1]);
        cudaEventRecord(e1, s[1]); cudaStreamWaitEvent(s[0], e1, 0);
      }
      cudaEventRecord(e1, s[0]); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%s pinned, %d stream(s): %.1f GB/s\n", (mode & 1) ? "write-combined" : "default", ns, n / best / 1e6);
  }
  return 0;
}
