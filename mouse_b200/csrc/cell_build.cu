// K1 (bond vector + midpoint, minimum image) and K2 (cell list: single-digit radix / counting
// sort by cell id, per-cell offsets, canonical in-cell order, payload gather).
//
// K1 restates lammpack_types.py:587-599 (np.remainder form) in un-fused FP64.  K2 has no
// reference counterpart - the reference is all-pairs (ordering_functions.py:94-115).
#include "common.cuh"

// np.remainder(a, b), b > 0: fmod then shift negatives by +b (NumPy npy_divmod); fmod is exact.
__device__ __forceinline__ double np_remainder_pos(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) {
    if (m < 0.0) m = __dadd_rn(m, b);
  } else {
    m = 0.0;
  }
  return m;
}

__device__ __forceinline__ void bond_axis(double p1, double p2, double L, double &v, double &c) {
  double x = __dsub_rn(p2, p1);
  double xc = __dmul_rn(__dadd_rn(p2, p1), 0.5);   // (p2 + p1) / 2.
  double half = __dmul_rn(L, 0.5);                 // L / 2.
  double xt = __dsub_rn(np_remainder_pos(__dadd_rn(x, half), L), half);
  v = xt;
  c = __dadd_rn(xc, __dmul_rn(__dsub_rn(xt, x), 0.5));
}

__global__ void __launch_bounds__(256) bond_build_kernel(const double *__restrict__ pos,
                                                         const int2 *__restrict__ bonds, int n,
                                                         double Lx, double Ly, double Lz,
                                                         double *__restrict__ vec, double *__restrict__ ctr) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int2 ab = __ldg(bonds + b);
  const double *p1 = pos + 3 * (size_t)ab.x;
  const double *p2 = pos + 3 * (size_t)ab.y;
  double v, c;
  bond_axis(__ldg(p1), __ldg(p2), Lx, v, c);
  vec[b] = v;
  ctr[b] = c;
  bond_axis(__ldg(p1 + 1), __ldg(p2 + 1), Ly, v, c);
  vec[(size_t)n + b] = v;
  ctr[(size_t)n + b] = c;
  bond_axis(__ldg(p1 + 2), __ldg(p2 + 2), Lz, v, c);
  vec[2 * (size_t)n + b] = v;
  ctr[2 * (size_t)n + b] = c;
}

// Wrapped copy of a coordinate into [0, L) and its cell index / in-cell offset.  The reference
// never wraps midpoints (lammpack_types.py:590-591); this copy is used for binning only.
__device__ __forceinline__ void cell_axis(double c, double L, double cw, int nc, int &k, double &xw) {
  double x = c - L * floor(c / L);
  if (x >= L) x -= L;
  if (!(x >= 0.0)) x = 0.0;
  k = (int)(x / cw);
  if (k >= nc) k = nc - 1;
  xw = x;
}

// One thread per item (bond midpoint or atom).  MODE 0: P2 (items = bonds, payload needs vec);
// MODE 1: RDF (items = atoms, xyz AoS [n][3], sel = type code < 0 means not selected).
template <int MODE>
__global__ void __launch_bounds__(256) cell_assign_kernel(const double *__restrict__ xyz,
                                                          const double *__restrict__ vec,
                                                          const uint8_t *__restrict__ nbr_mask,
                                                          const int32_t *__restrict__ type, int n, GridDev g,
                                                          int32_t *__restrict__ cell_of,
                                                          int32_t *__restrict__ rank_of,
                                                          int32_t *__restrict__ cell_count,
                                                          int64_t *__restrict__ counters) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  bool in_list;
  double cx, cy, cz;
  if (MODE == 0) {
    in_list = nbr_mask ? (nbr_mask[b] != 0) : true;
    double vx = vec[b], vy = vec[(size_t)n + b], vz = vec[2 * (size_t)n + b];
    if (in_list && vx == 0.0 && vy == 0.0 && vz == 0.0) {   // zero-length bond: never a reference
      atomicAdd((unsigned long long *)&counters[0], 1ULL);  // (:93,119), poisons as neighbour (:112)
      in_list = false;
    }
    cx = xyz[b];
    cy = xyz[(size_t)n + b];
    cz = xyz[2 * (size_t)n + b];
  } else {
    in_list = type[b] >= 0;
    cx = xyz[3 * (size_t)b];
    cy = xyz[3 * (size_t)b + 1];
    cz = xyz[3 * (size_t)b + 2];
  }
  if (!in_list) {
    cell_of[b] = -1;
    return;
  }
  int kx, ky, kz;
  double t;
  cell_axis(cx, g.box[0], g.cw[0], g.nc[0], kx, t);
  cell_axis(cy, g.box[1], g.cw[1], g.nc[1], ky, t);
  cell_axis(cz, g.box[2], g.cw[2], g.nc[2], kz, t);
  int cid = (kz * g.nc[1] + ky) * g.nc[0] + kx;
  cell_of[b] = cid;
  rank_of[b] = atomicAdd(&cell_count[cid], 1);
}

// ---- exclusive scan of cell_count[0..nc) -> cell_start, three small kernels ----------------
__global__ void __launch_bounds__(256) scan_reduce_kernel(const int32_t *__restrict__ in, int nc,
                                                          int32_t *__restrict__ block_sums) {
  __shared__ int s_warp[8];
  int base = blockIdx.x * MOUSE_SCAN_TILE;
  int sum = 0;
  for (int k = threadIdx.x; k < MOUSE_SCAN_TILE; k += 256) {
    int c = base + k;
    if (c < nc) sum += in[c];
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_warp[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) scan_top_kernel(int32_t *__restrict__ block_sums, int nb) {
  // one block; exclusive scan in place, chunks of 1024 with a running carry
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int k = base + threadIdx.x;
    int v = k < nb ? block_sums[k] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(MOUSE_FULL_MASK, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x];
      int xs = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(MOUSE_FULL_MASK, xs, o);
        if (threadIdx.x >= o) xs += y;
      }
      s_warp[threadIdx.x] = xs - w;   // exclusive warp offsets
    }
    __syncthreads();
    int carry = s_carry;
    int incl = x + s_warp[threadIdx.x >> 5];
    if (k < nb) block_sums[k] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + incl;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) scan_apply_kernel(const int32_t *__restrict__ in, int nc,
                                                         const int32_t *__restrict__ block_offs,
                                                         int32_t *__restrict__ out) {
  // each thread owns 8 consecutive entries of the tile
  __shared__ int s_warp[8];
  int base = blockIdx.x * MOUSE_SCAN_TILE + threadIdx.x * 8;
  int v[8];
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = base + k;
    v[k] = c < nc ? in[c] : 0;
    tsum += v[k];
  }
  int x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(MOUSE_FULL_MASK, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += s_warp[w];
  int run = block_offs[blockIdx.x] + woff + x - tsum;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = base + k;
    if (c < nc) out[c] = run;
    run += v[k];
  }
}

// ---- scatter payload into arrival order ------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) cell_scatter_kernel(const double *__restrict__ xyz,
                                                           const double *__restrict__ vec,
                                                           const int32_t *__restrict__ code,
                                                           const int32_t *__restrict__ code2,
                                                           const uint8_t *__restrict__ ref_mask, int n,
                                                           GridDev g, const int32_t *__restrict__ cell_of,
                                                           const int32_t *__restrict__ rank_of,
                                                           const int32_t *__restrict__ cell_start,
                                                           float4 *__restrict__ p_tmp,
                                                           double4 *__restrict__ u_tmp,
                                                           int32_t *__restrict__ idx_tmp) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int cid = cell_of[b];
  if (cid < 0) return;
  int dst = cell_start[cid] + rank_of[b];
  double cx, cy, cz;
  if (MODE == 0) {
    cx = xyz[b];
    cy = xyz[(size_t)n + b];
    cz = xyz[2 * (size_t)n + b];
  } else {
    cx = xyz[3 * (size_t)b];
    cy = xyz[3 * (size_t)b + 1];
    cz = xyz[3 * (size_t)b + 2];
  }
  int k;
  double xi, yi, zi;
  cell_axis(cx, g.box[0], g.cw[0], g.nc[0], k, xi);
  cell_axis(cy, g.box[1], g.cw[1], g.nc[1], k, yi);
  cell_axis(cz, g.box[2], g.cw[2], g.nc[2], k, zi);
  // FP32 copy of the wrapped coordinate, centred on the box (|x| <= L/2 keeps one more mantissa bit)
  float4 p;
  p.x = (float)(xi - 0.5 * g.box[0]);
  p.y = (float)(yi - 0.5 * g.box[1]);
  p.z = (float)(zi - 0.5 * g.box[2]);
  p.w = __int_as_float(code ? code[b] : 0);
  double4 u;
  if (MODE == 0) {
    double vx = vec[b], vy = vec[(size_t)n + b], vz = vec[2 * (size_t)n + b];
    double n2 = vx * vx + vy * vy + vz * vz;
    double inv = 1.0 / sqrt(n2);
    u = make_double4(vx * inv, vy * inv, vz * inv, n2);
  } else {
    u = make_double4(cx, cy, cz, code2 ? (double)code2[b] : 0.0);
  }
  p_tmp[dst] = p;
  u_tmp[dst] = u;
  bool is_ref = ref_mask ? (ref_mask[b] != 0) : true;
  idx_tmp[dst] = b | (is_ref ? (int32_t)0x80000000 : 0);
}

// ---- canonical order inside each cell: ascending original index (one warp per cell) ----------
__global__ void __launch_bounds__(256) cell_canon_kernel(int ncell, const int32_t *__restrict__ cell_start,
                                                         const float4 *__restrict__ p_tmp,
                                                         const double4 *__restrict__ u_tmp,
                                                         const int32_t *__restrict__ idx_tmp,
                                                         float4 *__restrict__ p_sorted,
                                                         double4 *__restrict__ u_sorted,
                                                         int32_t *__restrict__ idx_sorted,
                                                         int32_t *__restrict__ cell_sorted,
                                                         uint8_t *__restrict__ flag_sorted, int pack_index) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < ncell; c += warps) {
    int s = cell_start[c], e = cell_start[c + 1];
    int nc = e - s;
    if (nc <= 0) continue;
    if (nc <= 32) {
      int raw = lane < nc ? idx_tmp[s + lane] : 0x7fffffff;
      int mine = raw & 0x7fffffff;
      int rank = 0;
      for (int k = 0; k < nc; ++k) {
        int other = __shfl_sync(MOUSE_FULL_MASK, mine, k);
        rank += other < mine;
      }
      if (lane < nc) {
        int d = s + rank;
        p_sorted[d] = p_tmp[s + lane];
        double4 u = u_tmp[s + lane];
        if (pack_index) u.w = __hiloint2double(0, d);   // the sweep's candidate registers carry their own slot
        u_sorted[d] = u;
        idx_sorted[d] = mine;
        cell_sorted[d] = c;
        flag_sorted[d] = (uint8_t)((unsigned)raw >> 31);
      }
    } else {
      for (int m = lane; m < nc; m += 32) {
        int raw = idx_tmp[s + m];
        int mine = raw & 0x7fffffff;
        int rank = 0;
        for (int k = 0; k < nc; ++k) rank += (idx_tmp[s + k] & 0x7fffffff) < mine;
        int d = s + rank;
        p_sorted[d] = p_tmp[s + m];
        double4 u = u_tmp[s + m];
        if (pack_index) u.w = __hiloint2double(0, d);
        u_sorted[d] = u;
        idx_sorted[d] = mine;
        cell_sorted[d] = c;
        flag_sorted[d] = (uint8_t)((unsigned)raw >> 31);
      }
    }
  }
}

// ---- host-side launchers ------------------------------------------------------------------------
int mouse_launch_bond_build(const double *pos, const int32_t *bond_atoms, int64_t n, const double *box,
                            double *vec, double *ctr, cudaStream_t st) {
  if (n == 0) return MOUSE_OK;
  bond_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, (const int2 *)bond_atoms, (int)n, box[0],
                                                                  box[1], box[2], vec, ctr);
  MOUSE_LAUNCH_CHECK("bond_build_kernel");
  return MOUSE_OK;
}

// mode 0: xyz = ctr SoA [3][n], vec SoA, code = mol; mode 1: xyz = pos AoS [n][3], code = type, code2 = mol
int mouse_launch_cell_build(int mode, const double *xyz, const double *vec, const int32_t *code,
                            const int32_t *code2, const uint8_t *nbr_mask, const uint8_t *ref_mask, int64_t n, const mouse_grid_t *grid,
                            const Workspace &w, cudaStream_t st) {
  GridDev g = mouse_grid_dev(grid);
  int nc = grid->ncell + 1;
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.cell_count, 0, sizeof(int32_t) * (size_t)nc, st));
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters, 0, sizeof(int64_t) * 8, st));
  unsigned nb = (unsigned)((n + 255) / 256);
  if (n > 0) {
    if (mode == 0)
      cell_assign_kernel<0><<<nb, 256, 0, st>>>(xyz, vec, nbr_mask, nullptr, (int)n, g, w.cell_of, w.rank_of,
                                                w.cell_count, w.counters);
    else
      cell_assign_kernel<1><<<nb, 256, 0, st>>>(xyz, nullptr, nullptr, code, (int)n, g, w.cell_of, w.rank_of,
                                                w.cell_count, w.counters);
    MOUSE_LAUNCH_CHECK("cell_assign_kernel");
  }
  int tiles = (nc + MOUSE_SCAN_TILE - 1) / MOUSE_SCAN_TILE;
  scan_reduce_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks);
  MOUSE_LAUNCH_CHECK("scan_reduce_kernel");
  scan_top_kernel<<<1, 1024, 0, st>>>(w.scan_blocks, tiles);
  MOUSE_LAUNCH_CHECK("scan_top_kernel");
  scan_apply_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks, w.cell_start);
  MOUSE_LAUNCH_CHECK("scan_apply_kernel");
  if (n > 0) {
    if (mode == 0)
      cell_scatter_kernel<0><<<nb, 256, 0, st>>>(xyz, vec, code, nullptr, ref_mask, (int)n, g, w.cell_of, w.rank_of,
                                                 w.cell_start, w.p_tmp, w.u_tmp, w.idx_tmp);
    else
      cell_scatter_kernel<1><<<nb, 256, 0, st>>>(xyz, nullptr, code, code2, nullptr, (int)n, g, w.cell_of, w.rank_of,
                                                 w.cell_start, w.p_tmp, w.u_tmp, w.idx_tmp);
    MOUSE_LAUNCH_CHECK("cell_scatter_kernel");
    int blocks = (int)((((int64_t)grid->ncell * 32) + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cell_canon_kernel<<<blocks, 256, 0, st>>>(grid->ncell, w.cell_start, w.p_tmp, w.u_tmp, w.idx_tmp, w.p_sorted,
                                              w.u_sorted, w.idx_sorted, w.cell_sorted, w.flag_sorted, mode == 0);
    MOUSE_LAUNCH_CHECK("cell_canon_kernel");
  }
  return MOUSE_OK;
}
