// Microbenchmark: issue rate of the instruction classes of the half-shell sweep's inner loop on one SM
// (DFMA, packed FFMA2, scalar FFMA, integer ALU, and the loop's own mix) at the sweep's occupancy
// (1-warp CTAs, W warps per SM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_rate fp_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

template <int MODE>
__global__ void __launch_bounds__(32) rate_kernel(double *out, long long *cyc, int iters, double seed) {
  double a[8];
  float2 f[8];
  int q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x;
    f[i] = make_float2((float)(seed + i), (float)(seed - i));
    q[i] = (int)seed + i;
  }
  double b[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    b[i] = seed * (i + 2) + threadIdx.x;
    cc[i] = seed / (i + 3) - threadIdx.x;
  }
  const double m = seed * 0.5, c = seed * 0.25, e = seed * 0.125;
  const float2 mf = make_float2((float)m, (float)m), cf = make_float2((float)c, (float)c);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 16 DFMA
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
    } else if (MODE == 1) {   // 16 FFMA2
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __ffma2_rn(f[i], mf, cf);
    } else if (MODE == 2) {   // 16 scalar FFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i].x = __fmaf_rn(f[i].x, mf.x, cf.x);
        f[i].y = __fmaf_rn(f[i].y, mf.x, cf.x);
      }
    } else if (MODE == 3) {   // 8 DFMA + 8 FFMA2
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = fma(a[i], m, c);
        f[i] = __ffma2_rn(f[i], mf, cf);
      }
    } else if (MODE == 4) {   // 8 DFMA + 8 integer LOP/IADD
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a[i] = fma(a[i], m, c);
        q[i] = (q[i] ^ it) + i;
      }
    } else if (MODE == 5) {   // the loop's mix per slot pair: 10 DFMA + 7 FFMA2 + 11 integer/ALU
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(a[i], m, c);
      a[0] = fma(a[1], m, a[0]);
      a[2] = fma(a[3], m, a[2]);
#pragma unroll
      for (int i = 0; i < 7; ++i) f[i] = __ffma2_rn(f[i], mf, cf);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = (q[i] ^ it) + i;
      q[0] += q[1] & it;
      q[2] += q[3] & it;
      q[4] += q[5] & it;
    } else if (MODE == 6) {   // 16 integer
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = (q[i] ^ it) + i;
    } else if (MODE == 8) {   // 16 DFMA, three distinct register operands each (no operand reuse)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(b[i], cc[i], a[i]);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fma(cc[i], b[(i + 3) & 7], a[i]);
    } else if (MODE == 9) {   // the sweep's FP64 pattern for 4 candidates: 3-term dot, then two accumulations
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double d = fma(m, b[i], fma(c, cc[i], e * b[i + 4]));
        a[i] = fma(d, d, a[i]);
        a[i + 4] = fma(d, d, a[i + 4]);
      }
    } else if (MODE == 7) {   // 8 FFMA2 + 8 integer
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = __ffma2_rn(f[i], mf, cf);
        q[i] = (q[i] ^ it) + i;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + f[i].x + f[i].y + q[i] + b[i] + cc[i];
  out[blockIdx.x * 32 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_iter, double *out, long long *cyc, long long *h) {
  const int iters = 20000;
  for (int w : {4, 8, 12, 16, 24}) {
    const int blocks = 148 * w;
    rate_kernel<MODE><<<blocks, 32>>>(out, cyc, iters, 1.000001);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<blocks, 32>>>(out, cyc, iters, 1.000001);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ev_rate = (double)per_iter * iters * w / 4.0 / (ms * 1e-3 * 1.965e9);
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; ++i) mean += (double)h[i];
    mean /= blocks;
    // warp-instructions per clock per SM scheduler (4 per SM): w warps share 4 schedulers
    const double per_smsp = (double)per_iter * iters * w / 4.0 / mean;
    printf("%-24s warps/SM %2d: %.3f warp-instr/clk/SMSP by clock64, %.3f by events @1965 MHz (%.3f ms)\n", name, w, per_smsp, ev_rate, ms);
  }
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 148 * 24 * 32 * 8); cudaMalloc(&cyc, 148 * 24 * 8);
  long long *h = new long long[148 * 24];
  run<0>("DFMA", 16, out, cyc, h);
  run<8>("DFMA 3 distinct operands", 16, out, cyc, h);
  run<9>("DFMA dot+2 acc pattern", 20, out, cyc, h);
  run<1>("FFMA2 (packed)", 16, out, cyc, h);
  run<2>("FFMA (scalar)", 16, out, cyc, h);
  run<6>("integer LOP+IADD", 32, out, cyc, h);
  run<3>("DFMA + FFMA2 1:1", 16, out, cyc, h);
  run<4>("DFMA + 2 int 1:2", 24, out, cyc, h);
  run<7>("FFMA2 + 2 int 1:2", 24, out, cyc, h);
  run<5>("loop mix 10D+7F2+22I", 39, out, cyc, h);
  return 0;
}
