// K5: RDF pair histogram on the same cell list.  Replaces calculateRdfForReference
// (ordering_functions.py:66-88) summed over all reference atoms as in
// CalculatePairCorrelationFunctions (:162-173), counts accumulated onto zeros.
//
// v1: one thread per reference atom, all-FP64 distance in the reference's operation order
// (np.add(x, -1.*x_ref); dr + -1.*L*around(dr/L); sqrt(x^2+y^2+z^2)), NumPy's uniform-bin index
// rule (truncate, then +-1 correction against the edges table), shared-memory privatised
// histogram per block flushed with integer atomics (deterministic result).
#include "common.cuh"

struct RdfArgs {
  GridDev g;
  const float4 *P;       // .w = type code
  const double4 *U;      // absolute position (FP64) + molecule code in .w
  const int32_t *cell_start;
  const int32_t *cell_sorted;
  const double *edges;   // [nbin+1] device
  int nbin;
  int ntypes;
  int same_molecule;
  int use_smem;
  unsigned long long *hist;  // [npairs][nbin]
};

// numpy.histogram uniform-bin rule (NumPy 2.3 _histograms_impl.histogram)
__device__ __forceinline__ int np_bin(double d, double first, double last, int nbin, const double *edges) {
  if (!(d >= first) || !(d <= last)) return -1;
  double f = __dmul_rn(__ddiv_rn(__dsub_rn(d, first), __dsub_rn(last, first)), (double)nbin);
  int k = (int)f;
  if (k == nbin) k -= 1;
  if (d < edges[k]) k -= 1;
  if (d >= edges[k + 1] && k != nbin - 1) k += 1;
  return k;
}

__global__ void __launch_bounds__(128) rdf_sweep_exact_kernel(const RdfArgs a) {
  extern __shared__ unsigned int s_hist[];
  const int npairs = a.ntypes * (a.ntypes + 1) / 2;
  const int hsize = npairs * a.nbin;
  if (a.use_smem) {
    for (int k = threadIdx.x; k < hsize; k += blockDim.x) s_hist[k] = 0;
    __syncthreads();
  }
  const int n_list = a.cell_start[a.g.ncell];
  const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
  const double first = a.edges[0], last = a.edges[a.nbin];
  const double Lx = a.g.box[0], Ly = a.g.box[1], Lz = a.g.box[2];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_list; s += gridDim.x * blockDim.x) {
    const double4 pi = a.U[s];
    const int ti = __float_as_int(a.P[s].w);
    const int home = a.cell_sorted[s];
    const int cx = home % nx, cy = (home / nx) % ny, cz = home / (nx * ny);
    const int mx = nx >= 3 ? 3 : nx, my = ny >= 3 ? 3 : ny, mz = nz >= 3 ? 3 : nz;
    for (int kz = 0; kz < mz; ++kz) {
      const int zc = nz >= 3 ? (cz + kz - 1 + nz) % nz : kz;
      for (int ky = 0; ky < my; ++ky) {
        const int yc = ny >= 3 ? (cy + ky - 1 + ny) % ny : ky;
        for (int kx = 0; kx < mx; ++kx) {
          const int xc = nx >= 3 ? (cx + kx - 1 + nx) % nx : kx;
          const int cid = (zc * ny + yc) * nx + xc;
          const int ts = a.cell_start[cid], te = a.cell_start[cid + 1];
          for (int t = ts; t < te; ++t) {
            if (t == s) continue;
            const double4 pj = a.U[t];
            if (!a.same_molecule && pj.w == pi.w) continue;
            const double tx = mouse_trim(__dsub_rn(pj.x, pi.x), Lx);
            const double ty = mouse_trim(__dsub_rn(pj.y, pi.y), Ly);
            const double tz = mouse_trim(__dsub_rn(pj.z, pi.z), Lz);
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz)));
            if (d == 0.0) continue;  // np.ma.masked_equal(d, 0.) also drops coincident atoms (:84)
            const int k = np_bin(d, first, last, a.nbin, a.edges);
            if (k < 0) continue;
            const int tj = __float_as_int(a.P[t].w);
            const int ta = min(ti, tj), tb = max(ti, tj);
            const int slot = ta * a.ntypes - ta * (ta - 1) / 2 + (tb - ta);
            if (a.use_smem) atomicAdd(&s_hist[slot * a.nbin + k], 1u);
            else atomicAdd(&a.hist[(size_t)slot * a.nbin + k], 1ULL);
          }
        }
      }
    }
  }
  if (a.use_smem) {
    __syncthreads();
    for (int k = threadIdx.x; k < hsize; k += blockDim.x)
      if (s_hist[k]) atomicAdd(&a.hist[k], (unsigned long long)s_hist[k]);
  }
}

int mouse_launch_rdf_sweep(const mouse_grid_t *grid, uint32_t flags, int64_t n, int32_t ntypes,
                           const double *edges_dev, int32_t nbin, const Workspace &w, int64_t *hist,
                           cudaStream_t st) {
  RdfArgs a;
  a.g = mouse_grid_dev(grid);
  a.P = w.p_sorted;
  a.U = w.u_sorted;
  a.cell_start = w.cell_start;
  a.cell_sorted = w.cell_sorted;
  a.edges = edges_dev;
  a.nbin = nbin;
  a.ntypes = ntypes;
  a.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
  a.hist = (unsigned long long *)hist;
  const size_t hsize = (size_t)ntypes * (ntypes + 1) / 2 * (size_t)nbin;
  MOUSE_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int64_t) * hsize, st));
  if (n == 0) return MOUSE_OK;
  size_t smem = hsize * sizeof(unsigned int);
  a.use_smem = smem <= 96 * 1024;
  if (!a.use_smem) smem = 0;
  if (smem > 48 * 1024)
    MOUSE_CUDA_CHECK(cudaFuncSetAttribute(rdf_sweep_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
  int blocks = (int)((n + 127) / 128);
  if (blocks > 148 * 8) blocks = 148 * 8;
  rdf_sweep_exact_kernel<<<blocks, 128, smem, st>>>(a);
  MOUSE_LAUNCH_CHECK("rdf_sweep_exact_kernel");
  return MOUSE_OK;
}
