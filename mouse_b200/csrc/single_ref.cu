// One-reference kernels: the reference's per-call granularity.
//  * p2_single_ref_kernel  = calculateAveCosSqForReference (ordering_functions.py:91-119)
//  * rdf_single_ref_kernel = calculateRdfForReference      (ordering_functions.py:66-88)
// Both stream the caller's (9,N) / (6,N) float64 array once (HBM-bound), evaluate the
// reference's expression per column in its operation order, and reduce deterministically.
#include "common.cuh"

#define SR_THREADS 256
#define SR_MAX_BLOCKS 1024

struct SingleRef {
  double vx, vy, vz;     // reference bond vector      (bRef)
  double cx, cy, cz;     // reference midpoint / atom position
  double num;            // reference bond.num / atom.num as stored in row 0
  double mol;            // hash8(mol_id) as stored in row 2
  double box[3];
  double rc2;
  int use_cutoff, exclude_self, same_molecule;
};

__global__ void __launch_bounds__(SR_THREADS) p2_single_ref_kernel(const double *__restrict__ a, int n, SingleRef r,
                                                                   double *__restrict__ partials,
                                                                   int64_t *__restrict__ counters,
                                                                   double *__restrict__ out /*[2]: sum, count*/) {
  __shared__ double s_sum[SR_THREADS / 32];
  __shared__ long long s_cnt[SR_THREADS / 32];
  __shared__ bool s_last;
  const size_t N = (size_t)n;
  const double nr = __dadd_rn(__dadd_rn(__dmul_rn(r.vx, r.vx), __dmul_rn(r.vy, r.vy)), __dmul_rn(r.vz, r.vz));
  double acc = 0.0;
  long long cnt = 0;
  const int chunk = (n + gridDim.x - 1) / gridDim.x;
  const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  for (int j = lo + threadIdx.x; j < hi; j += SR_THREADS) {
    const double bx = a[3 * N + j], by = a[4 * N + j], bz = a[5 * N + j];
    double m = 1.0;
    if (r.use_cutoff) {
      double rsq = mouse_rsq_exact(r.cx, r.cy, r.cz, a[6 * N + j], a[7 * N + j], a[8 * N + j], r.box);
      if (!(rsq <= r.rc2)) m = 0.0;
    }
    if (r.exclude_self && a[j] == r.num) m = 0.0;
    if (!r.same_molecule && a[2 * N + j] == r.mol) m = 0.0;
    const double dot = mouse_dot_exact(r.vx, r.vy, r.vz, bx, by, bz);
    const double nj = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
    const double cs = __ddiv_rn(__ddiv_rn(__dmul_rn(m, __dmul_rn(dot, dot)), nr), nj);
    if (cs != 0.0) {   // np.ma.masked_equal(.., 0.): NaN is kept (and poisons the mean)
      acc += cs;
      cnt += 1;
    }
  }
  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(MOUSE_FULL_MASK, cnt, o);
  if ((threadIdx.x & 31) == 0) {
    s_sum[threadIdx.x >> 5] = acc;
    s_cnt[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    long long c = 0;
    for (int w = 0; w < SR_THREADS / 32; ++w) {
      t += s_sum[w];
      c += s_cnt[w];
    }
    partials[2 * blockIdx.x] = t;
    partials[2 * blockIdx.x + 1] = (double)c;
    __threadfence();
    s_last = atomicAdd((unsigned long long *)&counters[0], 1ULL) == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double t = 0.0, c = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      t += ((volatile double *)partials)[2 * b];
      c += ((volatile double *)partials)[2 * b + 1];
    }
    out[0] = t;
    out[1] = c;
    counters[0] = 0;
  }
}

// numpy.histogram uniform-bin rule (NumPy 2.3 _histograms_impl.histogram)
__device__ __forceinline__ int np_bin1(double d, double first, double last, int nbin, const double *edges) {
  if (!(d >= first) || !(d <= last)) return -1;
  double f = __dmul_rn(__ddiv_rn(__dsub_rn(d, first), __dsub_rn(last, first)), (double)nbin);
  int k = (int)f;
  if (k == nbin) k -= 1;
  if (d < edges[k]) k -= 1;
  if (d >= edges[k + 1] && k != nbin - 1) k += 1;
  return k;
}

__global__ void __launch_bounds__(SR_THREADS) rdf_single_ref_kernel(const double *__restrict__ a, int n, SingleRef r,
                                                                    const double *__restrict__ edges, int nbin,
                                                                    unsigned long long *__restrict__ hist
                                                                    /*[nbin + 1]: last = unmasked count*/) {
  extern __shared__ unsigned int s_h[];
  for (int k = threadIdx.x; k <= nbin; k += SR_THREADS) s_h[k] = 0;
  __syncthreads();
  const size_t N = (size_t)n;
  const double first = edges[0], last = edges[nbin];
  for (int j = blockIdx.x * SR_THREADS + threadIdx.x; j < n; j += gridDim.x * SR_THREADS) {
    double m = 1.0;
    if (r.exclude_self && a[j] == r.num) m = 0.0;
    if (!r.same_molecule && a[2 * N + j] == r.mol) m = 0.0;
    const double tx = mouse_trim(__dsub_rn(a[3 * N + j], r.cx), r.box[0]);
    const double ty = mouse_trim(__dsub_rn(a[4 * N + j], r.cy), r.box[1]);
    const double tz = mouse_trim(__dsub_rn(a[5 * N + j], r.cz), r.box[2]);
    const double d = __dmul_rn(m, sqrt(__dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz))));
    if (d == 0.0) continue;
    atomicAdd(&s_h[nbin], 1u);
    const int k = np_bin1(d, first, last, nbin, edges);
    if (k >= 0) atomicAdd(&s_h[k], 1u);
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= nbin; k += SR_THREADS)
    if (s_h[k]) atomicAdd(&hist[k], (unsigned long long)s_h[k]);
}

static void fill_ref(SingleRef &r, const double *ref, const double *box, double rcut, double rc2, uint32_t flags) {
  r.vx = ref[0]; r.vy = ref[1]; r.vz = ref[2];
  r.cx = ref[3]; r.cy = ref[4]; r.cz = ref[5];
  r.num = ref[6]; r.mol = ref[7];
  for (int k = 0; k < 3; ++k) r.box[k] = box[k];
  r.rc2 = rc2;
  r.use_cutoff = rcut >= 0.0;
  r.exclude_self = (flags & 8u) ? 0 : 1;
  r.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
}

extern "C" {

size_t mouse_single_ref_workspace_bytes(void) {
  return 256 + sizeof(double) * 2 * SR_MAX_BLOCKS + 256 + sizeof(double) * (MOUSE_MAX_BINS + 1) + 256;
}

int mouse_p2_single_ref(const double *np_bonds, int64_t n, const double *ref, const double *box, double rcut,
                        double rc2, uint32_t flags, void *ws, double *out, void *stream) {
  if (!np_bonds || !ref || !box || !ws || !out || n < 0 || n >= 2147483647LL) {
    mouse_set_error("mouse_p2_single_ref: bad argument");
    return MOUSE_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SingleRef r;
  fill_ref(r, ref, box, rcut, rc2, flags);
  int64_t *counters = (int64_t *)ws;
  double *partials = (double *)((char *)ws + 256);
  MOUSE_CUDA_CHECK(cudaMemsetAsync(counters, 0, 256, st));
  int blocks = (int)((n + SR_THREADS - 1) / SR_THREADS);
  blocks = blocks < 1 ? 1 : (blocks > SR_MAX_BLOCKS ? SR_MAX_BLOCKS : blocks);
  p2_single_ref_kernel<<<blocks, SR_THREADS, 0, st>>>(np_bonds, (int)n, r, partials, counters, out);
  MOUSE_LAUNCH_CHECK("p2_single_ref_kernel");
  return MOUSE_OK;
}

int mouse_rdf_single_ref(const double *np_atoms, int64_t n, const double *ref, const double *box, uint32_t flags,
                         const double *edges, int32_t nbin, void *ws, int64_t *hist, void *stream) {
  if (!np_atoms || !ref || !box || !ws || !hist || !edges || nbin < 1 || nbin > MOUSE_MAX_BINS || n < 0 ||
      n >= 2147483647LL) {
    mouse_set_error("mouse_rdf_single_ref: bad argument");
    return MOUSE_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  SingleRef r;
  double ref8[8] = {0, 0, 0, ref[0], ref[1], ref[2], ref[3], ref[4]};
  fill_ref(r, ref8, box, 0.0, 0.0, flags);
  double *edges_dev = (double *)((char *)ws + 256 + sizeof(double) * 2 * SR_MAX_BLOCKS + 256);
  MOUSE_CUDA_CHECK(cudaMemcpyAsync(edges_dev, edges, sizeof(double) * (size_t)(nbin + 1), cudaMemcpyHostToDevice, st));
  MOUSE_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int64_t) * (size_t)(nbin + 1), st));
  int blocks = (int)((n + SR_THREADS - 1) / SR_THREADS);
  blocks = blocks < 1 ? 1 : (blocks > 148 * 4 ? 148 * 4 : blocks);
  rdf_single_ref_kernel<<<blocks, SR_THREADS, sizeof(unsigned int) * (size_t)(nbin + 1), st>>>(
      np_atoms, (int)n, r, edges_dev, nbin, (unsigned long long *)hist);
  MOUSE_LAUNCH_CHECK("rdf_single_ref_kernel");
  return MOUSE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// P2 of every bond against ONE director (CalculateCrystallinityParameter, histo branch, ordering_functions.py:263-267):
//   s = 1.5 * (b.x*d.x + b.y*d.y + b.z*d.z)**2 / b.length_sq() / d.length_sq() - 0.5     (left to right, un-fused)
// vec is the SoA [3][n] output of K1; bonds with mask == 0 get NaN and are not histogrammed.  The histogram uses the
// reference's updateHistogram rule (lammpack_misc.py:185-189): int(nbins*(v-min)/(max-min)), negative wraps, >= nbins
// is counted in *over (the reference raises IndexError).
__global__ void __launch_bounds__(256) director_p2_kernel(const double *__restrict__ vec, const uint8_t *__restrict__ mask,
                                                          int n, double dx, double dy, double dz,
                                                          double *__restrict__ out_s, unsigned long long *__restrict__ hist,
                                                          unsigned long long *__restrict__ over, double smin, double smax,
                                                          int nbins) {
  extern __shared__ unsigned int s_hist[];   // nbins + 1
  for (int k = threadIdx.x; k <= nbins; k += blockDim.x) s_hist[k] = 0;
  __syncthreads();
  const size_t N = (size_t)n;
  const double ld = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
    double s = __longlong_as_double(0x7ff8000000000000LL);
    if (!mask || mask[b]) {
      const double bx = vec[b], by = vec[N + b], bz = vec[2 * N + b];
      const double dot = __dadd_rn(__dadd_rn(__dmul_rn(bx, dx), __dmul_rn(by, dy)), __dmul_rn(bz, dz));
      const double lb = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
      s = __dsub_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(1.5, __dmul_rn(dot, dot)), lb), ld), 0.5);
      const double f = __ddiv_rn(__dmul_rn((double)nbins, __dsub_rn(s, smin)), __dsub_rn(smax, smin));
      int bin = nbins;
      if (f == f && fabs(f) < 2.0e9) {
        const int t = (int)f;   // truncation toward zero, like Python int()
        if (t >= 0 && t < nbins) bin = t;
        else if (t < 0 && t >= -nbins) bin = nbins + t;
      }
      atomicAdd(&s_hist[bin], 1u);
    }
    out_s[b] = s;
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= nbins; k += blockDim.x)
    if (s_hist[k]) atomicAdd(k < nbins ? &hist[k] : over, (unsigned long long)s_hist[k]);
}

int mouse_launch_director_p2(const double *vec, const uint8_t *mask, int64_t n, const double *director, double *out_s,
                             int64_t *hist, double smin, double smax, int32_t nbins, cudaStream_t st) {
  MOUSE_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int64_t) * (size_t)(nbins + 1), st));
  if (n == 0) return MOUSE_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  director_p2_kernel<<<blocks, 256, sizeof(unsigned int) * (size_t)(nbins + 1), st>>>(
      vec, mask, (int)n, director[0], director[1], director[2], out_s, (unsigned long long *)hist,
      (unsigned long long *)hist + nbins, smin, smax, nbins);
  MOUSE_LAUNCH_CHECK("director_p2_kernel");
  return MOUSE_OK;
}
