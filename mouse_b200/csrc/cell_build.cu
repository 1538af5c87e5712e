// K1 (bond vector + midpoint, minimum image) and K2 (cell list: single-digit radix / counting
// sort by cell id, per-cell offsets, canonical in-cell order, payload gather).
//
// K1 restates lammpack_types.py:587-599 (np.remainder form) in un-fused FP64.  K2 has no
// reference counterpart - the reference is all-pairs (ordering_functions.py:94-115).
#include <string.h>

#include "half_shell.cuh"

// Threads per block of the per-item HBM-side kernels.  Small blocks fit into the registers the persistent sweep of
// ANOTHER frame leaves free on an SM (10 CTAs x 32 threads x 168 registers = 53 760 of 65 536), so the build of frame
// k+1 really runs next to the sweep of frame k; alone on the GPU they reach the same occupancy as large blocks.
#ifndef MOUSE_ITEM_BLOCK
#define MOUSE_ITEM_BLOCK 128
#endif

// np.remainder(a, b), b > 0: fmod then shift negatives by +b (NumPy npy_divmod); fmod is exact.
// For |a| < 2 b (every bond of a wrapped or mildly unwrapped configuration) fmod needs no library call:
// fmod(a, b) = a for |a| < b, and a -+ b for b <= |a| < 2 b, where the subtraction is exact (Sterbenz).
__device__ __forceinline__ double np_remainder_pos(double a, double b) {
  double m;
  const double aa = fabs(a);
  if (aa < b) m = a;
  else if (aa < __dadd_rn(b, b)) m = a < 0.0 ? __dadd_rn(a, b) : __dsub_rn(a, b);
  else m = fmod(a, b);
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

// Triclinic box (build-defined, see include/mouse_b200.h): the same np.remainder step axis by axis from z to x; a shift
// by n lattice vectors along an axis also removes n times their tilt components from the lower axes.  tilt = {xy, xz, yz}
__device__ __forceinline__ void bond_tri(const double *p1, const double *p2, const double *box, const double *tilt,
                                         double (&v)[3], double (&c)[3]) {
  double x[3], xc[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double q1 = __ldg(p1 + a), q2 = __ldg(p2 + a);
    x[a] = __dsub_rn(q2, q1);
    xc[a] = __dmul_rn(__dadd_rn(q2, q1), 0.5);
  }
  const double hz = __dmul_rn(box[2], 0.5), hy = __dmul_rn(box[1], 0.5), hx = __dmul_rn(box[0], 0.5);
  v[2] = __dsub_rn(np_remainder_pos(__dadd_rn(x[2], hz), box[2]), hz);
  const double nz = rint(__ddiv_rn(__dsub_rn(x[2], v[2]), box[2]));
  const double y1 = __dsub_rn(x[1], __dmul_rn(tilt[2], nz));
  const double x1 = __dsub_rn(x[0], __dmul_rn(tilt[1], nz));
  v[1] = __dsub_rn(np_remainder_pos(__dadd_rn(y1, hy), box[1]), hy);
  const double ny = rint(__ddiv_rn(__dsub_rn(y1, v[1]), box[1]));
  const double x2 = __dsub_rn(x1, __dmul_rn(tilt[0], ny));
  v[0] = __dsub_rn(np_remainder_pos(__dadd_rn(x2, hx), box[0]), hx);
#pragma unroll
  for (int a = 0; a < 3; ++a) c[a] = __dadd_rn(xc[a], __dmul_rn(__dsub_rn(v[a], x[a]), 0.5));
}

// bond vector + midpoint of one bond for either kind of box
__device__ __forceinline__ void bond_geom(const double *p1, const double *p2, const GridDev &g, double (&v)[3],
                                          double (&c)[3]) {
  if (g.tri) {
    bond_tri(p1, p2, g.box, g.tilt, v, c);
  } else {
    bond_axis(__ldg(p1), __ldg(p2), g.box[0], v[0], c[0]);
    bond_axis(__ldg(p1 + 1), __ldg(p2 + 1), g.box[1], v[1], c[1]);
    bond_axis(__ldg(p1 + 2), __ldg(p2 + 2), g.box[2], v[2], c[2]);
  }
}

__global__ void __launch_bounds__(256) bond_build_kernel(const double *__restrict__ pos,
                                                         const int2 *__restrict__ bonds, int n, GridDev g,
                                                         double *__restrict__ vec, double *__restrict__ ctr) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int2 ab = __ldg(bonds + b);
  double v[3], c[3];
  bond_geom(pos + 3 * (size_t)ab.x, pos + 3 * (size_t)ab.y, g, v, c);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    vec[a * (size_t)n + b] = v[a];
    ctr[a * (size_t)n + b] = c[a];
  }
}

// Wrapped copy of a coordinate into [0, L) and its cell index / in-cell offset.  The reference
// never wraps midpoints (lammpack_types.py:590-591); this copy is used for binning only.
__device__ __forceinline__ void cell_axis(double c, double L, double cw, int nc, int &k, double &xw) {
  // c - L * floor(c / L); the division is only needed outside [-L, 2 L)
  double x;
  if (c >= 0.0 && c < L) x = c;
  else if (c >= -L && c < 0.0) x = c + L;
  else if (c >= L && c < L + L) x = c - L;
  else x = c - L * floor(c / L);
  if (x >= L) x -= L;
  if (!(x >= 0.0)) x = 0.0;
  k = (int)(x / cw);
  if (k >= nc) k = nc - 1;
  xw = x;
}

// Cell of a point and its wrapped copy for either kind of box.  Triclinic: the point is expressed along the lattice
// directions (fractional coordinate times edge length), wrapped and binned there, and mapped back to Cartesian; the
// FP32 payload is the wrapped Cartesian position minus the centre of the cell H (1/2, 1/2, 1/2).  With tilt = 0 all
// of this reduces to the orthorhombic expressions bit for bit.
__device__ __forceinline__ int cell_of_point(double cx, double cy, double cz, const GridDev &g, double &xw, double &yw,
                                             double &zw) {
  int kx, ky, kz;
  if (g.tri) {
    const double fz = cz / g.box[2];
    const double uy = cy - g.tilt[2] * fz;
    const double ux = cx - g.tilt[0] * (uy / g.box[1]) - g.tilt[1] * fz;
    double uxw, uyw, uzw;
    cell_axis(ux, g.box[0], g.cw[0], g.nc[0], kx, uxw);
    cell_axis(uy, g.box[1], g.cw[1], g.nc[1], ky, uyw);
    cell_axis(cz, g.box[2], g.cw[2], g.nc[2], kz, uzw);
    xw = uxw + g.tilt[0] * (uyw / g.box[1]) + g.tilt[1] * (uzw / g.box[2]);
    yw = uyw + g.tilt[2] * (uzw / g.box[2]);
    zw = uzw;
  } else {
    cell_axis(cx, g.box[0], g.cw[0], g.nc[0], kx, xw);
    cell_axis(cy, g.box[1], g.cw[1], g.nc[1], ky, yw);
    cell_axis(cz, g.box[2], g.cw[2], g.nc[2], kz, zw);
  }
  return (kz * g.nc[1] + ky) * g.nc[0] + kx;
}

// payload of an item, written once in ORIGINAL order (16 + 32 contiguous bytes: the canonical gather then costs two
// sectors per item).  P = FP32 copy of the wrapped coordinate, centred on the box (|x| <= L/2 keeps one more
// mantissa bit) + code bits; U = FP64 unit bond vector (P2) / absolute position + molecule code (RDF).
__device__ __forceinline__ float4 payload_p(double xw, double yw, double zw, const GridDev &g, int code) {
  float4 p;
  p.x = (float)(xw - g.ctr_off[0]);
  p.y = (float)(yw - g.ctr_off[1]);
  p.z = (float)(zw - g.ctr_off[2]);
  p.w = __int_as_float(code);
  return p;
}
__device__ __forceinline__ double4 payload_u_bond(double vx, double vy, double vz) {
  double n2 = vx * vx + vy * vy + vz * vz;
  double inv = 1.0 / sqrt(n2);
  return make_double4(vx * inv, vy * inv, vz * inv, 0.0);
}

// One thread per item (bond midpoint or atom).  MODE 0: P2 (items = bonds, payload needs vec);
// MODE 1: RDF (items = atoms, xyz AoS [n][3], sel = type code < 0 means not selected).
template <int MODE>
__global__ void __launch_bounds__(256) cell_assign_kernel(const double *__restrict__ xyz,
                                                          const double *__restrict__ vec,
                                                          const uint8_t *__restrict__ nbr_mask,
                                                          const int32_t *__restrict__ type,
                                                          const int32_t *__restrict__ code2, int n, GridDev g,
                                                          int32_t *__restrict__ cell_of,
                                                          int32_t *__restrict__ rank_of,
                                                          int32_t *__restrict__ cell_count,
                                                          int64_t *__restrict__ counters,
                                                          Rec *__restrict__ rec_tmp) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  bool in_list;
  double cx, cy, cz;
  double vx = 0.0, vy = 0.0, vz = 0.0;
  if (MODE == 0) {
    in_list = nbr_mask ? (nbr_mask[b] != 0) : true;
    vx = vec[b], vy = vec[(size_t)n + b], vz = vec[2 * (size_t)n + b];
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
  double xw, yw, zw;
  const int cid = cell_of_point(cx, cy, cz, g, xw, yw, zw);
  cell_of[b] = cid;
  MOUSE_ASSERT(cid >= 0 && cid < g.ncell);
  rank_of[b] = atomicAdd(&cell_count[cid], 1);
  Rec r;
  r.p = payload_p(xw, yw, zw, g, type ? type[b] : 0);
  r.u = MODE == 0 ? payload_u_bond(vx, vy, vz) : make_double4(cx, cy, cz, code2 ? (double)code2[b] : 0.0);
  rec_tmp[b] = r;
}

// ---- heaviest cells first -------------------------------------------------------------------
// The half-shell sweeps hand out home cells through a ticket counter; a crowded cell (cost ~ occupancy^2) taken late
// leaves the other warps idle at the end of the kernel.  The item rows are therefore written in eight occupancy
// classes, heaviest first (longest-processing-time-first scheduling): class of a cell = its occupancy relative to the
// mean, in steps of half the mean; empty cells last.  counters[16 + k] = item rows of class k (histogram taken by the scan),
// counters[24 + k] = rows of class k handed out so far; the scan leaves every cell's row in place of its occupancy.
#define MOUSE_NCLASS 8
// item rows of a cell: one, or one per window of `window` items (P2 sweep); window == 0: always one
__device__ __forceinline__ int cell_rows(int count, int window) {
  return (window > 0 && count > window) ? (count + window - 1) / window : 1;
}
__device__ __forceinline__ int cell_class(int count, float inv_half_mean) {
  if (count <= 0) return MOUSE_NCLASS - 1;
  const int q = (int)((float)count * inv_half_mean);       // occupancy in units of half the mean
  return q >= 6 ? 0 : 6 - q;
}

// Rows of the 8 consecutive cells a thread of the scan owns (v = their occupancies, first cell `base`): the block
// reserves one contiguous range per class (one global atomic per class and tile), its cells take consecutive rows
// inside it (a crowded cell takes one row per window).  The FIRST row overwrites the cell's occupancy in place (dead
// after the scan).  All threads of the block call.
__device__ __forceinline__ void lpt_rows(const int (&v)[8], int base, int ncells, int32_t *__restrict__ rows,
                                         int64_t *counters, float inv_half_mean, int window) {
  __shared__ int s_n[MOUSE_NCLASS], s_first[MOUSE_NCLASS];
  if (threadIdx.x < MOUSE_NCLASS) s_n[threadIdx.x] = 0;
  __syncthreads();
  int cls[8], rk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    cls[k] = cell_class(v[k], inv_half_mean);
    rk[k] = base + k < ncells ? atomicAdd(&s_n[cls[k]], cell_rows(v[k], window)) : 0;
  }
  __syncthreads();
  if (threadIdx.x < MOUSE_NCLASS) {
    int first = 0;      // rows of the heavier classes come first
    for (int j = 0; j < (int)threadIdx.x; ++j) first += (int)counters[16 + j];
    const int mine = s_n[threadIdx.x];
    const int tile = mine ? (int)atomicAdd((unsigned long long *)&counters[24 + threadIdx.x], (unsigned long long)mine) : 0;
    s_first[threadIdx.x] = first + tile;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (base + k < ncells) rows[base + k] = s_first[cls[k]] + rk[k];
}

// ---- exclusive scan of cell_count[0..nc) -> cell_start, three small kernels ----------------
__global__ void __launch_bounds__(256) scan_reduce_kernel(const int32_t *__restrict__ in, int nc,
                                                          int32_t *__restrict__ block_sums, int64_t *counters,
                                                          float inv_half_mean, int window) {
  __shared__ int s_warp[8];
  __shared__ int s_hist[MOUSE_NCLASS];
  if (threadIdx.x < MOUSE_NCLASS) s_hist[threadIdx.x] = 0;
  __syncthreads();
  int base = blockIdx.x * MOUSE_SCAN_TILE;
  int sum = 0;
  for (int k = threadIdx.x; k < MOUSE_SCAN_TILE; k += 256) {
    int c = base + k;
    if (c < nc - 1) {      // (the last entry of the array is the unused total slot, not a cell)
      const int v = in[c];
      sum += v;
      atomicAdd(&s_hist[cell_class(v, inv_half_mean)], cell_rows(v, window));
    }
  }
  __syncthreads();
  if (counters && threadIdx.x < MOUSE_NCLASS && s_hist[threadIdx.x])
    atomicAdd((unsigned long long *)&counters[16 + threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
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

__global__ void __launch_bounds__(256) scan_apply_kernel(int32_t *in, int nc,
                                                         const int32_t *__restrict__ block_offs,
                                                         int32_t *__restrict__ out, int64_t *counters,
                                                         float inv_half_mean, int window) {
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
  lpt_rows(v, base, nc - 1, in, counters, inv_half_mean, window);    // in[c] becomes the cell's first row in the item table
}

// ---- exclusive scan, second kernel of two: every block sums the tile totals in front of it (a few dozen values)
// and scans its own tile
__global__ void __launch_bounds__(256) scan_apply2_kernel(int32_t *in, int nc,
                                                          const int32_t *__restrict__ tile_sums,
                                                          int32_t *__restrict__ out, int64_t *counters,
                                                          float inv_half_mean, int window) {
  __shared__ int s_warp[8];
  __shared__ int s_off;
  // offset of this tile = sum of the totals of all tiles before it
  int part = 0;
  for (int k = threadIdx.x; k < (int)blockIdx.x; k += 256) part += tile_sums[k];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_warp[w];
    s_off = t;
  }
  __syncthreads();
  const int off = s_off;
  __syncthreads();
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
  int run = off + woff + x - tsum;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    int c = base + k;
    if (c < nc) out[c] = run;
    run += v[k];
  }
  lpt_rows(v, base, nc - 1, in, counters, inv_half_mean, window);    // in[c] becomes the cell's first row in the item table
}

// ---- scatter the item index into its cell (arrival order); the payload moves once, in cell_place_kernel -----
__global__ void __launch_bounds__(MOUSE_ITEM_BLOCK) cell_scatter_idx_kernel(const uint8_t *__restrict__ ref_mask, int n,
                                                               const int32_t *__restrict__ cell_of,
                                                               const int32_t *__restrict__ rank_of,
                                                               const int32_t *__restrict__ cell_start,
                                                               int32_t *__restrict__ idx_tmp) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int cid = cell_of[b];
  if (cid < 0) return;
  bool is_ref = ref_mask ? (ref_mask[b] != 0) : true;
  MOUSE_ASSERT(rank_of[b] >= 0 && cell_start[cid] + rank_of[b] < cell_start[cid + 1] && cell_start[cid + 1] <= n);
  idx_tmp[cell_start[cid] + rank_of[b]] = b | (is_ref ? (int32_t)0x80000000 : 0);
}

// ---- canonical order inside each cell (ascending original index) + payload gather: ONE THREAD PER ITEM -----------
// The items of a cell arrive in atomic order; ranking them by original index makes the sorted arrays - and with them
// every floating-point sum of the sweeps - a deterministic function of the input.  Thread p owns the item at arrival
// position p: its rank = items of its cell with a smaller original index (the cell's arrival list is read by all the
// threads of the cell: broadcast loads), then it gathers its 48-byte record (2 sectors) and writes the sorted arrays.
// Every lane works whatever the occupancy of the cells (a warp per cell keeps 14 of 32 lanes busy at 1M bonds and
// ranks one huge cell - rmax == 0, no cutoff - serially: seconds for 100 000 bonds).
template <int MODE>
__global__ void __launch_bounds__(MOUSE_ITEM_BLOCK) cell_place_kernel(int ncell, const int32_t *__restrict__ cell_start,
                                                                      const int32_t *__restrict__ idx_tmp,
                                                                      const Rec *__restrict__ rec_tmp, int32_t *slot_of,
                                                                      Rec *__restrict__ rec_sorted,
                                                                      int32_t *__restrict__ idx_sorted,
                                                                      int32_t *__restrict__ cell_sorted,
                                                                      uint8_t *__restrict__ flag_sorted,
                                                                      long long *__restrict__ acc_fix,
                                                                      int32_t *__restrict__ acc_cnt,
                                                                      uint32_t *__restrict__ cell_bbox) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < cell_start[ncell];   // items in the list
  const unsigned act = __ballot_sync(MOUSE_FULL_MASK, live);     // the lanes that meet again at the box reduction
  if (!live) return;
  const int raw = idx_tmp[p];
  const int mine = raw & 0x7fffffff;
  const int c = slot_of[mine];               // still the item's cell id (this thread replaces it below)
  const int s = cell_start[c], e = cell_start[c + 1];
  int rank = 0;
  for (int k = s; k < e; ++k) rank += (idx_tmp[k] & 0x7fffffff) < mine;
  const int d = s + rank;
  MOUSE_ASSERT(c >= 0 && c < ncell && s <= p && p < e && d >= s && d < e);
  Rec r = rec_tmp[mine];
  // .w carries the sorted slot itself: the sweep's register-resident candidates know where they live
  if (MODE == 0) r.u.w = __hiloint2double(0, d);
  rec_sorted[d] = r;
  // inverse permutation (the cell id of the item is dead by now): sorted slot | reference bit << 30, so that
  // the fused finish + reduce walks the outputs in original order (coalesced stores, gathers that hit L2)
  slot_of[mine] = d | (int)(((unsigned)raw >> 31) << 30);
  idx_sorted[d] = mine;
  cell_sorted[d] = c;
  flag_sorted[d] = (uint8_t)((unsigned)raw >> 31);
  acc_fix[d] = 0;      // the half-shell sweep's accumulators start (and are left) at zero
  acc_cnt[d] = 0;
  if (MODE == 0 && cell_bbox) {
    // bounding box of the cell's items (the P2 sweep drops candidates farther than the cutoff from it): the threads of
    // a cell are neighbours, so the lanes of the warp that share a cell reduce their keys (REDUX) and one of them
    // folds the result into the cell's box with six integer atomics
    const unsigned peers = __match_any_sync(act, c);
    uint32_t k[6];
    k[0] = mouse_float_key(r.p.x), k[1] = mouse_float_key(r.p.y), k[2] = mouse_float_key(r.p.z);
#pragma unroll
    for (int a = 0; a < 3; ++a) k[3 + a] = ~k[a];
#pragma unroll
    for (int a = 0; a < 6; ++a) k[a] = __reduce_max_sync(peers, k[a]);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) {
#pragma unroll
      for (int a = 0; a < 6; ++a) atomicMax(&cell_bbox[(size_t)c * 6 + a], k[a]);
    }
  }
}

// ---- the item rows of the half-shell sweeps (half_shell.cuh): ONE THREAD PER CELL ------------------------------
// [0..9] first sorted slot of range L, [10..19] its length, [20..29] its offset in the candidate concatenation,
// [30] size | boundary << 30, [31] centre z, [32..37] bounding box of the cell's items (P2: the sweep drops candidates
// farther than the cutoff from it), [38..39] centre x, y.  The row index comes from the scan (heaviest cells first).
template <int MODE>
__global__ void __launch_bounds__(128) cell_items_kernel(int ncell, GridDev g, const int32_t *__restrict__ cell_start,
                                                         const uint32_t *__restrict__ cell_bbox,
                                                         const int32_t *__restrict__ cell_row,
                                                         int32_t *__restrict__ items, int window) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int s = cell_start[c], nc = cell_start[c + 1] - s;
  int4 *row = (int4 *)(items + (size_t)cell_row[c] * HS_ITEM_INTS);
  if (nc <= 0) {       // an empty cell only needs its size (the sweeps skip it)
    row[7] = make_int4(0, 0, 0, 0);
    row[10] = make_int4(0, 0, 0, 0);
    return;
  }
  int cx, cy, cz;
  hs_cell_coords(g, c, cx, cy, cz);
  int v[40];
  int off = 0;
#pragma unroll
  for (int L = 0; L < 10; ++L) {
    int rb, lo, hi, src = 0, len = 0;
    hs_range_cells(g, cx, cy, cz, L, rb, lo, hi);
    if (lo <= hi) {
      src = cell_start[rb + lo];
      len = cell_start[rb + hi + 1] - src;
    }
    v[L] = src;
    v[10 + L] = len;
    v[20 + L] = off;
    off += len;
  }
  const int nx = g.nc[0], ny = g.nc[1], nz = g.nc[2];
  const int boundary = (cx == 0) | (cx == nx - 1) | (cy == 0) | (cy == ny - 1) | (cz == nz - 1);
  // cell centre along the lattice directions, then Cartesian (the tilt ratios are 0 for an orthorhombic box)
  const float ux = ((float)cx + 0.5f) * g.cwf[0] - 0.5f * g.boxf[0];
  const float uy = ((float)cy + 0.5f) * g.cwf[1] - 0.5f * g.boxf[1];
  const float uz = ((float)cz + 0.5f) * g.cwf[2] - 0.5f * g.boxf[2];
  v[30] = nc | (boundary << 30);
  v[31] = __float_as_int(uz);
  v[38] = __float_as_int(ux + g.tratio[0] * uy + g.tratio[1] * uz);
  v[39] = __float_as_int(uy + g.tratio[2] * uz);
#pragma unroll
  for (int a = 0; a < 3; ++a) {      // upper corner: max of the keys; lower corner: max of the complemented keys
    v[35 + a] = MODE == 0 ? __float_as_int(mouse_key_float(cell_bbox[(size_t)c * 6 + a])) : 0;
    v[32 + a] = MODE == 0 ? __float_as_int(mouse_key_float(~cell_bbox[(size_t)c * 6 + 3 + a])) : 0;
  }
#pragma unroll
  for (int q = 0; q < 10; ++q) row[q] = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  row[10] = make_int4(0, 0, 0, 0);
  // a crowded cell has one row per window of references (consecutive rows: the same ranges, their own window index), so
  // that its windows are separate work items - several warps share the cell instead of one warp sweeping it alone
  const int nw = cell_rows(nc, window);
  for (int w = 1; w < nw; ++w) {
    int4 *rw = row + (size_t)w * (HS_ITEM_INTS / 4);
#pragma unroll
    for (int q = 0; q < 10; ++q) rw[q] = make_int4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    rw[10] = make_int4(w, 0, 0, 0);
  }
}

// ---- K1 + cell assignment + output initialisation in one pass (mouse_p2_frame) -------------------------------
__global__ void __launch_bounds__(MOUSE_ITEM_BLOCK) bond_assign_kernel(const double *__restrict__ pos, const int2 *__restrict__ bonds,
                                                          const int32_t *__restrict__ mol,
                                                          const uint8_t *__restrict__ nbr_mask, int n, GridDev g,
                                                          double *__restrict__ vec, double *__restrict__ ctr,
                                                          int32_t *__restrict__ cell_of, int32_t *__restrict__ rank_of,
                                                          int32_t *__restrict__ cell_count,
                                                          int64_t *__restrict__ counters, Rec *__restrict__ rec_tmp,
                                                          double *__restrict__ out_sum,
                                                          int32_t *__restrict__ out_cnt, double *__restrict__ out_mean,
                                                          int init_listed) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  int2 ab = __ldg(bonds + b);
  double vv[3], cc[3];
  bond_geom(pos + 3 * (size_t)ab.x, pos + 3 * (size_t)ab.y, g, vv, cc);
  const double vx = vv[0], vy = vv[1], vz = vv[2], cx = cc[0], cy = cc[1], cz = cc[2];
  vec[b] = vx;
  vec[(size_t)n + b] = vy;
  vec[2 * (size_t)n + b] = vz;
  ctr[b] = cx;
  ctr[(size_t)n + b] = cy;
  ctr[2 * (size_t)n + b] = cz;
  bool in_list = nbr_mask ? (nbr_mask[b] != 0) : true;
  if (in_list && vx == 0.0 && vy == 0.0 && vz == 0.0) {   // zero-length bond: never a reference
    atomicAdd((unsigned long long *)&counters[0], 1ULL);  // (:93,119), poisons as neighbour (:112)
    in_list = false;
  }
  if (!in_list || init_listed) {
    // bonds outside the cell list never reach the fused K4: their outputs are final here (listed bonds are written
    // by p2_finish_reduce_kernel; the all-FP64 sweep + plain reduce path needs every output initialised)
    out_sum[b] = 0.0;
    out_cnt[b] = 0;
    out_mean[b] = __longlong_as_double(0x7ff8000000000000LL);
  }
  if (!in_list) {
    cell_of[b] = -1;
    return;
  }
  double xw, yw, zw;
  const int cid = cell_of_point(cx, cy, cz, g, xw, yw, zw);
  cell_of[b] = cid;
  MOUSE_ASSERT(cid >= 0 && cid < g.ncell);
  rank_of[b] = atomicAdd(&cell_count[cid], 1);
  Rec r;
  r.p = payload_p(xw, yw, zw, g, mol ? mol[b] : 0);
  r.u = payload_u_bond(vx, vy, vz);
  rec_tmp[b] = r;
}

// ---- host-side launchers ------------------------------------------------------------------------
int mouse_launch_bond_build(const double *pos, const int32_t *bond_atoms, int64_t n, const double *box,
                            const double *tilt, double *vec, double *ctr, cudaStream_t st) {
  if (n == 0) return MOUSE_OK;
  mouse_grid_t gh;
  memset(&gh, 0, sizeof(gh));
  for (int a = 0; a < 3; ++a) {
    gh.box[a] = box[a];
    gh.cell_w[a] = box[a];
    gh.ncell_axis[a] = 1;
    gh.tilt[a] = tilt ? tilt[a] : 0.0;
  }
  gh.ncell = 1;
  gh.triclinic = tilt ? 1 : 0;
  bond_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pos, (const int2 *)bond_atoms, (int)n,
                                                                  mouse_grid_dev(&gh), vec, ctr);
  MOUSE_LAUNCH_CHECK("bond_build_kernel");
  return MOUSE_OK;
}

// exclusive scan of cell_count -> cell_start, then index scatter and canonical gather (shared by all cell builds)
static int cell_build_tail(int mode, const uint8_t *ref_mask, int64_t n, const mouse_grid_t *grid, const GridDev &g,
                           const Workspace &w, cudaStream_t st) {
  int nc = grid->ncell + 1;
  int tiles = (nc + MOUSE_SCAN_TILE - 1) / MOUSE_SCAN_TILE;
  // occupancy classes of the item rows (heaviest cells first): in units of half the mean occupancy
  const float inv_half_mean = (float)(2.0 * (double)grid->ncell / (double)(n > 0 ? n : 1));
  const int window = mode == 0 ? MOUSE_WINDOW : 0;      // P2: one item row per window of references; RDF: one per cell
  if (tiles <= 4096) {   // the apply kernel sums the tile totals in front of its tile itself
    scan_reduce_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks, w.counters, inv_half_mean, window);
    MOUSE_LAUNCH_CHECK("scan_reduce_kernel");
    scan_apply2_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks, w.cell_start, w.counters, inv_half_mean, window);
    MOUSE_LAUNCH_CHECK("scan_apply2_kernel");
  } else {
    scan_reduce_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks, w.counters, inv_half_mean, window);
    MOUSE_LAUNCH_CHECK("scan_reduce_kernel");
    scan_top_kernel<<<1, 1024, 0, st>>>(w.scan_blocks, tiles);
    MOUSE_LAUNCH_CHECK("scan_top_kernel");
    scan_apply_kernel<<<tiles, 256, 0, st>>>(w.cell_count, nc, w.scan_blocks, w.cell_start, w.counters, inv_half_mean, window);
    MOUSE_LAUNCH_CHECK("scan_apply_kernel");
  }
  if (n > 0) {
    unsigned nb = (unsigned)((n + MOUSE_ITEM_BLOCK - 1) / MOUSE_ITEM_BLOCK);
    cell_scatter_idx_kernel<<<nb, MOUSE_ITEM_BLOCK, 0, st>>>(ref_mask, (int)n, w.cell_of, w.rank_of, w.cell_start, w.idx_tmp);
    MOUSE_LAUNCH_CHECK("cell_scatter_idx_kernel");
    if (mode == 0)
      cell_place_kernel<0><<<nb, MOUSE_ITEM_BLOCK, 0, st>>>(grid->ncell, w.cell_start, w.idx_tmp, w.rec_tmp, w.cell_of, w.rec_sorted,
                                                            w.idx_sorted, w.cell_sorted, w.flag_sorted, w.acc_fix, w.acc_cnt, grid->fast ? w.cell_bbox : nullptr);
    else
      cell_place_kernel<1><<<nb, MOUSE_ITEM_BLOCK, 0, st>>>(grid->ncell, w.cell_start, w.idx_tmp, w.rec_tmp, w.cell_of, w.rec_sorted,
                                                            w.idx_sorted, w.cell_sorted, w.flag_sorted, w.acc_fix, w.acc_cnt, grid->fast ? w.cell_bbox : nullptr);
    MOUSE_LAUNCH_CHECK("cell_place_kernel");
  }
  if (grid->fast) {      // (also for an empty list: the sweeps read every cell's row)
    const unsigned cb = (unsigned)((grid->ncell + 127) / 128);
    if (mode == 0)
      cell_items_kernel<0><<<cb, 128, 0, st>>>(grid->ncell, g, w.cell_start, w.cell_bbox, w.cell_count, w.items, window);
    else
      cell_items_kernel<1><<<cb, 128, 0, st>>>(grid->ncell, g, w.cell_start, w.cell_bbox, w.cell_count, w.items, window);
    MOUSE_LAUNCH_CHECK("cell_items_kernel");
  }
  return MOUSE_OK;
}

// mode 0: xyz = ctr SoA [3][n], vec SoA, code = mol; mode 1: xyz = pos AoS [n][3], code = type, code2 = mol
int mouse_launch_cell_build(int mode, const double *xyz, const double *vec, const int32_t *code,
                            const int32_t *code2, const uint8_t *nbr_mask, const uint8_t *ref_mask, int64_t n, const mouse_grid_t *grid,
                            const Workspace &w, cudaStream_t st) {
  GridDev g = mouse_grid_dev(grid);
  // counters sit right in front of cell_count: one memset clears both
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters, 0, mouse_clear_bytes(w, grid->ncell), st));
  unsigned nb = (unsigned)((n + 255) / 256);
  if (n > 0) {
    if (mode == 0)
      cell_assign_kernel<0><<<nb, 256, 0, st>>>(xyz, vec, nbr_mask, code, nullptr, (int)n, g, w.cell_of, w.rank_of,
                                                w.cell_count, w.counters, w.rec_tmp);
    else
      cell_assign_kernel<1><<<nb, 256, 0, st>>>(xyz, nullptr, nullptr, code, code2, (int)n, g, w.cell_of, w.rank_of,
                                                w.cell_count, w.counters, w.rec_tmp);
    MOUSE_LAUNCH_CHECK("cell_assign_kernel");
  }
  return cell_build_tail(mode, ref_mask, n, grid, g, w, st);
}

// K1 + K2 of mouse_p2_frame: bond vectors / midpoints, cell list, outputs initialised (sum 0, count 0, mean NaN)
int mouse_launch_frame_build(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                             const uint8_t *ref_mask, int64_t n, const mouse_grid_t *grid, const Workspace &w,
                             double *vec, double *ctr, double *out_sum, int32_t *out_cnt, double *out_mean,
                             int init_listed, cudaStream_t st) {
  GridDev g = mouse_grid_dev(grid);
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters, 0, mouse_clear_bytes(w, grid->ncell), st));
  if (n > 0) {
    bond_assign_kernel<<<(unsigned)((n + MOUSE_ITEM_BLOCK - 1) / MOUSE_ITEM_BLOCK), MOUSE_ITEM_BLOCK, 0, st>>>(pos, (const int2 *)bond_atoms, mol, nbr_mask, (int)n,
                                                                     g, vec, ctr, w.cell_of, w.rank_of, w.cell_count,
                                                                     w.counters, w.rec_tmp, out_sum, out_cnt,
                                                                     out_mean, init_listed);
    MOUSE_LAUNCH_CHECK("bond_assign_kernel");
  }
  return cell_build_tail(0, ref_mask, n, grid, g, w, st);
}
