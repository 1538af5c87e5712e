// Microbenchmark: throughput of coalesced RED.ADD (s64 / s32) into an L2-resident array, the access pattern of the
// half-shell sweep's per-candidate flush.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red red.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(32) red_kernel(unsigned long long *fix, int *cnt, int n, int ncell, int slots, int mode) {
  const int lane = threadIdx.x;
  for (int cell = blockIdx.x; cell < ncell; cell += gridDim.x) {
    unsigned h = (unsigned)cell * 2654435761u;
    int base = (int)(h % (unsigned)(n - 32 * slots));
    if (mode & 4) base = (int)(((long long)cell * 13) % (n - 32 * slots));   // near-sequential
    for (int c = 0; c < slots; ++c) {
      int j = base + 32 * c + lane;
      if (mode & 1) atomicAdd(&fix[j], (unsigned long long)(j + cell));
      if (mode & 2) atomicAdd(&cnt[j], 1);
    }
  }
}
int main() {
  const int n = 1000000;
  unsigned long long *fix; int *cnt;
  cudaMalloc(&fix, n * 8); cudaMalloc(&cnt, n * 4);
  cudaMemset(fix, 0, n * 8); cudaMemset(cnt, 0, n * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ncell = 90000, slots = 8;
  for (int mode : {1, 2, 3, 5, 6, 7}) {
    for (int blocks : {148 * 4, 148 * 12, 148 * 32}) {
      red_kernel<<<blocks, 32>>>(fix, cnt, n, ncell, slots, mode);
      cudaEventRecord(e0);
      for (int r = 0; r < 5; ++r) red_kernel<<<blocks, 32>>>(fix, cnt, n, ncell, slots, mode);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
      double lane_reds = (double)ncell * slots * 32 * ((mode & 1 ? 1 : 0) + (mode & 2 ? 1 : 0));
      printf("mode %d blocks %5d: %.1f us  %.1f G lane-RED/s  (err %s)\n", mode, blocks, ms * 1e3, lane_reds / ms / 1e6,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
