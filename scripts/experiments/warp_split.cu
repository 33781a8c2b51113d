// Does a warp always reconverge after a divergent, latency-bound per-lane loop on sm_100a?  (profiles/README.md, "the lost
// records of round 1".)  Every lane walks a pointer chain of a random length through a table much larger than L2; most lanes
// have nothing to do (like padding slots of a token list).  After the loop: __syncwarp(), then the active mask is sampled.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_split warp_split.cu ; run on a B200.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) walk_kernel(const uint32_t* tab, uint32_t mask, const uint32_t* work, uint32_t n, unsigned long long* out) {
  __shared__ uint32_t s_cnt;
  const uint32_t lane = threadIdx.x & 31;
  for (uint32_t base = blockIdx.x * 256u; base < n; base += gridDim.x * 256u) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint32_t i = base + threadIdx.x;
    uint32_t w = i < n ? work[i] : 0u;
    uint32_t steps = w & 127u, p = w >> 7;
    bool hit = false;
    if (steps) {
      for (uint32_t k = 0; k < steps; k++) { p = tab[p & mask]; if ((p & 0xFFFu) == 0x123u) break; }
      hit = (p & 7u) == 3u;
    }
    __syncwarp();
    const uint32_t am = __activemask();
    if (am != 0xFFFFFFFFu && (am & (0u - am)) == (1u << lane)) atomicAdd(&out[0], 1ULL);  // one count per sub-group
    if (hit) atomicAdd(&out[1], 1ULL);
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
    if (lane == 0 && bal) atomicAdd(&out[2], (unsigned long long)__popc(bal));
    if (hit) atomicAdd(&s_cnt, 1u);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(&out[3], (unsigned long long)s_cnt);
    __syncthreads();
  }
}

int main(int argc, char** argv) {
  const uint32_t tab_words = 1u << 28;  // 1 GiB: every step is a DRAM access
  const uint32_t n = argc > 1 ? (uint32_t)atol(argv[1]) : 128u << 20;
  const int density = argc > 2 ? atoi(argv[2]) : 8;  // per cent of slots that walk
  std::vector<uint32_t> tab(tab_words), work(n);
  uint64_t s = 88172645463325252ULL;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
  for (auto& x : tab) x = rnd();
  for (auto& x : work) x = (rnd() % 100 < (uint32_t)density) ? ((rnd() << 7) | (1 + rnd() % 100)) : 0u;
  uint32_t *d_tab, *d_work; unsigned long long* d_out;
  cudaMalloc(&d_tab, (size_t)tab_words * 4); cudaMalloc(&d_work, (size_t)n * 4); cudaMalloc(&d_out, 64);
  cudaMemcpy(d_tab, tab.data(), (size_t)tab_words * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_work, work.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
  for (int run = 0; run < 6; run++) {
    cudaMemset(d_out, 0, 64);
    walk_kernel<<<148 * 8, 256>>>(d_tab, tab_words - 1, d_work, n, d_out);
    unsigned long long h[4];
    cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost);
    printf("run %d: %s  partial sub-groups %llu, hits per thread %llu, per ballot %llu, per block %llu\n", run, cudaGetErrorString(cudaGetLastError()), h[0], h[1], h[2], h[3]);
  }
  return 0;
}
