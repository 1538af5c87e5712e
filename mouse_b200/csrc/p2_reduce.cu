// K4: fused per-vector finalisation and whole-system reduction.
// Replaces ordering_functions.py:202-206, 221-224 (S = 1.5 * sum_i m_i / i_s - 0.5, a mean of
// per-vector means) and updateHistogram (lammpack_misc.py:185-189).
#include "common.cuh"

#define RED_THREADS 256
#ifndef FIN_THREADS
#define FIN_THREADS 128    // fused finish + reduce: small blocks fit next to another frame's sweep (see cell_build.cu)
#endif

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
    __syncthreads();   // every thread's out-of-range count is in counters[3] before thread 0 takes the ticket
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

// ---- fused finish + reduce (mouse_p2_frame, fast sweep): the half-shell sweep's fixed-point accumulators (per
// sorted slot) are converted, written to out_sum / out_cnt / out_mean in original bond order (the next frame build zeroes them again),
// and reduced to the totals in the same pass.  Deterministic: fixed slot ranges per block, fixed trees.
struct FinishArgs {
  const long long *acc_fix;
  const int32_t *acc_cnt;
  const int32_t *slot_of;   // [n] original index -> sorted slot | reference bit << 30, or -1 (not in the cell list)
  int n;
  int chunk;
  double *out_sum;
  int32_t *out_cnt;
  double *out_mean;
  double *partials;
  int64_t *counters;
  mouse_p2_totals_t *totals;
};

__global__ void __launch_bounds__(FIN_THREADS) p2_finish_reduce_kernel(const FinishArgs a) {
  __shared__ double s_sum[FIN_THREADS / 32];
  __shared__ long long s_refs[FIN_THREADS / 32], s_pairs[FIN_THREADS / 32];
  __shared__ bool s_last;
  const bool overflow = ((volatile int64_t *)a.counters)[4] != 0;   // the exact sweep wrote out_sum / out_cnt
  const long long nzero = a.counters[0];
  const int lo = blockIdx.x * a.chunk;
  const int hi = min(a.n, lo + a.chunk);
  const long long *__restrict__ acc_fix = a.acc_fix;
  const int32_t *__restrict__ acc_cnt = a.acc_cnt;
  const int32_t *__restrict__ slot_of = a.slot_of;
  double acc = 0.0;
  long long refs = 0, pairs = 0;
  // Original bond order: the stores are coalesced, the accumulators are gathered through the inverse permutation
  // (they were last touched by the sweep's atomics, i.e. they sit in L2).  Four bonds per thread and trip, all
  // loads up front (memory-level parallelism).
  for (int b0 = lo + threadIdx.x; b0 < hi; b0 += 4 * FIN_THREADS) {
    int v[4];
    long long fix[4];
    int c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int b = b0 + q * FIN_THREADS;
      v[q] = b < hi ? slot_of[b] : -1;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int s = v[q] & 0x3fffffff;
      fix[q] = v[q] >= 0 ? acc_fix[s] : 0;
      c[q] = v[q] >= 0 ? acc_cnt[s] : 0;
    }
    // (the accumulators are NOT re-zeroed here: every frame build zeroes them in cell_place_kernel)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int b = b0 + q * FIN_THREADS;
      if (v[q] < 0) continue;              // not in the cell list: the build wrote 0 / 0 / NaN
      const double nan = __longlong_as_double(0x7ff8000000000000LL);
      if (!((v[q] >> 30) & 1)) {           // listed, but not a reference: 0 / 0 / NaN
        a.out_sum[b] = 0.0;
        a.out_cnt[b] = 0;
        a.out_mean[b] = nan;
        continue;
      }
      double sum;
      int cc = c[q];
      if (overflow) {
        sum = a.out_sum[b];
        cc = a.out_cnt[b];
      } else {
        sum = (double)fix[q] * 7.105427357601002e-15;   // 2^-47
        if (nzero) {   // zero-length bonds poison every mean and add to every count (ordering_functions.py:112)
          sum = __longlong_as_double(0x7ff8000000000000LL);
          cc += (int)nzero;
        }
        a.out_sum[b] = sum;
        a.out_cnt[b] = cc;
      }
      double m = nan;
      if (cc > 0) {
        m = __ddiv_rn(sum, (double)cc);
        acc += m;
        refs += 1;
        pairs += cc;
      }
      a.out_mean[b] = m;
    }
  }
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
  if (threadIdx.x == 0) {
    double t = 0.0;
    long long r = 0, p = 0;
    for (int w = 0; w < FIN_THREADS / 32; ++w) {
      t += s_sum[w];
      r += s_refs[w];
      p += s_pairs[w];
    }
    a.partials[3 * blockIdx.x] = t;
    a.partials[3 * blockIdx.x + 1] = (double)r;   // exact: < 2^53
    a.partials[3 * blockIdx.x + 2] = (double)p;
    __threadfence();
    unsigned long long ticket = atomicAdd((unsigned long long *)&a.counters[5], 1ULL);
    s_last = (ticket == (unsigned long long)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double t = 0.0, r = 0.0, p = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += FIN_THREADS) {
      t += ((volatile double *)a.partials)[3 * b];
      r += ((volatile double *)a.partials)[3 * b + 1];
      p += ((volatile double *)a.partials)[3 * b + 2];
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
      for (int w = 0; w < FIN_THREADS / 32; ++w) {
        tt += s_sum[w];
        rr += s_refs[w];
        pp += s_pairs[w];
      }
      a.totals->sum_mean = tt;
      a.totals->n_refs = rr;
      a.totals->n_pairs = pp;
      a.totals->n_zero_len = ((volatile int64_t *)a.counters)[0];
      a.totals->n_band = ((volatile int64_t *)a.counters)[1];
      a.totals->n_hist_over = 0;
    }
  }
}

int mouse_launch_p2_finish_reduce(int64_t n, const mouse_grid_t *grid, const Workspace &w, double *out_sum,
                                  int32_t *out_cnt, double *out_mean, mouse_p2_totals_t *totals, cudaStream_t st) {
  FinishArgs a;
  a.acc_fix = w.acc_fix;
  a.acc_cnt = w.acc_cnt;
  a.slot_of = w.cell_of;
  a.n = (int)n;
  (void)grid;
  int blocks = (int)((n + FIN_THREADS - 1) / FIN_THREADS);
  if (blocks > MOUSE_REDUCE_BLOCKS) blocks = MOUSE_REDUCE_BLOCKS;
  if (blocks < 1) blocks = 1;
  int chunk = (int)((n + blocks - 1) / blocks);
  chunk = (chunk + FIN_THREADS - 1) / FIN_THREADS * FIN_THREADS;
  if (chunk < FIN_THREADS) chunk = FIN_THREADS;
  a.chunk = chunk;
  a.out_sum = out_sum;
  a.out_cnt = out_cnt;
  a.out_mean = out_mean;
  a.partials = w.partials;
  a.counters = w.counters;
  a.totals = totals;
  p2_finish_reduce_kernel<<<blocks, FIN_THREADS, 0, st>>>(a);
  MOUSE_LAUNCH_CHECK("p2_finish_reduce_kernel");
  return MOUSE_OK;
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

// ---- per-atom colouring (--out-pdb): ordering_functions.py:207-220.  Per atom, the mean of 1.5 * m - 0.5 over its
// incident bonds that produced a value, summed in ascending bond order (a per-atom list gathered through a CSR
// adjacency: the reference appends in bond order and np.mean of fewer than 8 values adds them sequentially), then
// the integer of the type string: code = int(mean * 100) >= 0 -> "C%02d", or -(int(mean * -100)) - 1 < 0 -> "Cm%02d";
// INT_MIN = no bond with a value (the atom keeps its type).
__global__ void __launch_bounds__(256) atom_colour_kernel(const double *__restrict__ mean, const int32_t *__restrict__ cnt,
                                                          const int64_t *__restrict__ row_ptr,
                                                          const int32_t *__restrict__ bond_of, int64_t n_atoms, int direct,
                                                          int32_t *__restrict__ code) {
  const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  double sum = 0.0;
  int k = 0;
  for (int64_t e = row_ptr[a]; e < row_ptr[a + 1]; ++e) {
    const int b = bond_of[e];
    if (cnt[b] > 0) {
      sum = __dadd_rn(sum, direct ? mean[b] : __dsub_rn(__dmul_rn(1.5, mean[b]), 0.5));
      k += 1;
    }
  }
  int c = (int)0x80000000;
  if (k > 0) {
    const double m = __ddiv_rn(sum, (double)k);
    c = m >= 0.0 ? (int)__dmul_rn(m, 100.0) : -(int)__dmul_rn(m, -100.0) - 1;
  }
  code[a] = c;
}

int mouse_launch_atom_colour(const double *mean, const int32_t *cnt, const int64_t *row_ptr, const int32_t *bond_of,
                             int64_t n_atoms, int32_t direct, int32_t *code, cudaStream_t st) {
  if (n_atoms == 0) return MOUSE_OK;
  atom_colour_kernel<<<(unsigned)((n_atoms + 255) / 256), 256, 0, st>>>(mean, cnt, row_ptr, bond_of, n_atoms, direct, code);
  MOUSE_LAUNCH_CHECK("atom_colour_kernel");
  return MOUSE_OK;
}
