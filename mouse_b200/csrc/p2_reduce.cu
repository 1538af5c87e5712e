// K4: fused per-vector finalisation and whole-system reduction.
// Replaces ordering_functions.py:202-206, 221-224 (S = 1.5 * sum_i m_i / i_s - 0.5, a mean of
// per-vector means) and updateHistogram (lammpack_misc.py:185-189).
#include "common.cuh"

#define RED_THREADS 256

__global__ void __launch_bounds__(RED_THREADS) p2_reduce_kernel(
    const double *__restrict__ sum, const int32_t *__restrict__ cnt, int n, int chunk,
    double *__restrict__ out_mean, double *__restrict__ partials, int64_t *__restrict__ counters,
    mouse_p2_totals_t *__restrict__ totals, int64_t *__restrict__ hist, double smin, double smax, int nbins) {
  extern __shared__ unsigned int s_hist[];  // nbins + 1 (last = out of range) when hist != NULL
  __shared__ double s_sum[RED_THREADS / 32];
  __shared__ long long s_refs[RED_THREADS / 32], s_pairs[RED_THREADS / 32];
  __shared__ bool s_last;
  if (hist) {
    for (int k = threadIdx.x; k <= nbins; k += RED_THREADS) s_hist[k] = 0;
    __syncthreads();
  }
  const int lo = blockIdx.x * chunk;
  const int hi = min(n, lo + chunk);
  double acc = 0.0;
  long long refs = 0, pairs = 0;
  for (int b = lo + threadIdx.x; b < hi; b += RED_THREADS) {
    const int c = cnt[b];
    double m = __longlong_as_double(0x7ff8000000000000LL);
    if (c > 0) {
      m = __ddiv_rn(sum[b], (double)c);
      acc += m;
      refs += 1;
      pairs += c;
      if (hist) {
        // int(nbins * (v - min) / (max - min)) with Python list indexing (negative wraps)
        const double v = __dsub_rn(__dmul_rn(1.5, m), 0.5);
        const double f = __ddiv_rn(__dmul_rn((double)nbins, __dsub_rn(v, smin)), __dsub_rn(smax, smin));
        int bin = nbins;
        if (f == f && fabs(f) < 2.0e9) {
          const int t = (int)f;  // truncation toward zero, like Python int()
          if (t >= 0 && t < nbins) bin = t;
          else if (t < 0 && t >= -nbins) bin = nbins + t;
        }
        atomicAdd(&s_hist[bin], 1u);
      }
    }
    out_mean[b] = m;
  }
  // fixed-order block reduction (deterministic)
  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    refs += __shfl_xor_sync(MOUSE_FULL_MASK, refs, o);
    pairs += __shfl_xor_sync(MOUSE_FULL_MASK, pairs, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = acc;
    s_refs[threadIdx.x >> 5] = refs;
    s_pairs[threadIdx.x >> 5] = pairs;
  }
  __syncthreads();
  if (hist) {
    for (int k = threadIdx.x; k <= nbins; k += RED_THREADS)
      if (s_hist[k]) {
        if (k < nbins) atomicAdd((unsigned long long *)&hist[k], (unsigned long long)s_hist[k]);
        else atomicAdd((unsigned long long *)&counters[3], (unsigned long long)s_hist[k]);
      }
  }
  if (threadIdx.x == 0) {
    double t = 0.0;
    long long r = 0, p = 0;
    for (int w = 0; w < RED_THREADS / 32; ++w) {
      t += s_sum[w];
      r += s_refs[w];
      p += s_pairs[w];
    }
    partials[3 * blockIdx.x] = t;
    partials[3 * blockIdx.x + 1] = (double)r;   // exact: < 2^53
    partials[3 * blockIdx.x + 2] = (double)p;
    __threadfence();
    unsigned long long ticket = atomicAdd((unsigned long long *)&counters[2], 1ULL);
    s_last = (ticket == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    // the last block sums the block partials: fixed assignment and fixed tree -> deterministic
    __threadfence();
    double t = 0.0, r = 0.0, p = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += RED_THREADS) {
      t += ((volatile double *)partials)[3 * b];
      r += ((volatile double *)partials)[3 * b + 1];
      p += ((volatile double *)partials)[3 * b + 2];
    }
    t = warp_sum(t);
    r = warp_sum(r);
    p = warp_sum(p);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
      s_sum[threadIdx.x >> 5] = t;
      s_refs[threadIdx.x >> 5] = (long long)r;
      s_pairs[threadIdx.x >> 5] = (long long)p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tt = 0.0;
      long long rr = 0, pp = 0;
      for (int w = 0; w < RED_THREADS / 32; ++w) {
        tt += s_sum[w];
        rr += s_refs[w];
        pp += s_pairs[w];
      }
      totals->sum_mean = tt;
      totals->n_refs = rr;
      totals->n_pairs = pp;
      totals->n_zero_len = ((volatile int64_t *)counters)[0];
      totals->n_band = ((volatile int64_t *)counters)[1];
      totals->n_hist_over = ((volatile int64_t *)counters)[3];
      counters[2] = 0;
    }
  }
}

int mouse_launch_p2_reduce(const double *sum, const int32_t *cnt, int64_t n, const Workspace &w, double *out_mean,
                           mouse_p2_totals_t *totals, int64_t *hist, double smin, double smax, int32_t nbins,
                           cudaStream_t st) {
  int blocks = (int)((n + RED_THREADS - 1) / RED_THREADS);
  if (blocks > MOUSE_REDUCE_BLOCKS) blocks = MOUSE_REDUCE_BLOCKS;
  if (blocks < 1) blocks = 1;
  int chunk = (int)((n + blocks - 1) / blocks);
  chunk = (chunk + RED_THREADS - 1) / RED_THREADS * RED_THREADS;
  if (chunk < RED_THREADS) chunk = RED_THREADS;
  size_t smem = 0;
  if (hist) {
    if (nbins < 1 || nbins > 8192) {
      mouse_set_error("nbins=%d out of range [1, 8192]", nbins);
      return MOUSE_EINVAL;
    }
    smem = sizeof(unsigned int) * (size_t)(nbins + 1);
    MOUSE_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int64_t) * (size_t)nbins, st));
  }
  // n_hist_over lives in counters[3]; reset it (counters[0..1] belong to the cell build / sweep)
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters + 2, 0, sizeof(int64_t) * 2, st));
  p2_reduce_kernel<<<blocks, RED_THREADS, smem, st>>>(sum, cnt, (int)n, chunk, out_mean, w.partials, w.counters,
                                                      totals, hist, smin, smax, nbins);
  MOUSE_LAUNCH_CHECK("p2_reduce_kernel");
  return MOUSE_OK;
}
