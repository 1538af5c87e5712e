// Internal helpers shared by the kernels of libmouse_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mouse_b200.h"

#define MOUSE_FULL_MASK 0xffffffffu

// Checked build (make EXTRA=-DMOUSE_CHECKED): device-side assertions on the indices the kernels compute themselves
// (landing-zone and queue slots, sorted-slot and cell indices, histogram bins, item rows).  compute-sanitizer is not
// available on the GPU pool this was developed on; the GPU test suite run against the checked library is the bounds
// evidence kept under profiles/.  In the normal build the macro expands to nothing.
#ifdef MOUSE_CHECKED
#include <assert.h>
#define MOUSE_ASSERT(cond) assert(cond)
#else
#define MOUSE_ASSERT(cond) ((void)0)
#endif

// ------------------------------------------------------------------ error plumbing
void mouse_set_error(const char *fmt, ...);
int mouse_num_sms();      // SMs of the current device (p2_sweep.cu: per-device table of immutable facts)

#define MOUSE_CUDA_CHECK(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      mouse_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,   \
                      __LINE__);                                                          \
      return MOUSE_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define MOUSE_LAUNCH_CHECK(name)                                                          \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      mouse_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));           \
      return MOUSE_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

// One item of a cell list as the sweeps consume it: 48 bytes, 16-byte aligned, so that one 1-D TMA bulk copy moves a
// whole range of items and a 48-byte stride is conflict-free for 128-bit shared-memory loads.
//   p = FP32 (wrapped coordinate - L/2) x, y, z + molecule (P2) / type (RDF) code bits in .w
//   u = P2: FP64 unit bond vector, .w = bit pattern of the item's own sorted slot;  RDF: absolute position, .w = molecule code
struct __align__(16) Rec {
  float4 p;
  double4 u;
};
static_assert(sizeof(Rec) == 48, "Rec must be 48 bytes");

// ------------------------------------------------------------------ workspace layout
// Everything the cell list and the sweeps need, carved out of one caller-provided buffer.
struct Workspace {
  // per cell
  int32_t *cell_count;   // [ncell+1]   items per cell (atomic histogram), last entry unused; after the scan: the cell's row in `items`
  uint32_t *cell_bbox;   // [ncell][6]  bounding box of the cell's items as order-preserving integer keys, reduced with
                         //             atomicMax: [0..2] upper corner, [3..5] lower corner (complemented keys); 0 = empty
  int32_t *cell_start;   // [ncell+1]   exclusive scan; cell_start[ncell] = items in the list
  int32_t *scan_blocks;  // [scan_nblocks+1]
  int32_t *items;        // [ncell + n / 32 + 1][44] item rows of the half-shell sweeps (ranges of the 14 cells, see half_shell.cuh):
                         //               one row per cell (RDF) or per window of MOUSE_WINDOW references of a cell (P2)
  // per item, original order
  int32_t *cell_of;      // [n] cell id or -1 (not in the list); after the build: sorted slot | reference bit << 30
                         //     (inverse permutation, still -1 for items outside the list)
  int32_t *rank_of;      // [n] arrival rank inside the cell (atomic order, canonicalised later)
  // per item, cell-sorted (ascending original index inside a cell)
  Rec *rec_tmp;          // [n] payload in ORIGINAL order (built by the assign kernel, gathered by the canon kernel);
                         //     reused after the build as the fast sweep's pair list (6 int2 entries per item)
  int32_t *idx_tmp;      // [n] original index (| ref bit), arrival order
  Rec *rec_sorted;       // [n] the payload in cell-sorted order
  int32_t *idx_sorted;
  int32_t *cell_sorted;  // [n] cell id of each sorted slot
  uint8_t *flag_sorted;  // [n] bit0 = acts as reference
  // half-shell sweep accumulators per sorted slot (adjacent: zeroed with one memset)
  long long *acc_fix;    // [n] sum of cos^2, 2^-47 fixed point (integer atomics: exact, order independent)
  int32_t *acc_cnt;      // [n] counted neighbours
  // scalars
  int64_t *counters;     // [32]: 0 zero-length bonds in the neighbour set, 1 exact re-decisions, 2 home-cell ticket /
                         //       reduce ticket, 3 pair-list length / histogram overflow, 4 pair-list overflow flag,
                         //       5 ticket and 6 histogram overflow of the fused finish + reduce
  double *partials;      // [3 * MOUSE_REDUCE_BLOCKS] block partials of the fused reduction
  double *edges;         // [MOUSE_MAX_BINS + 1] device copy of the RDF bin edges
  double *thr;           // [MOUSE_MAX_BINS + 2] the edges as exact thresholds on d^2 (fast RDF sweep)
  double rdf_first;      // (host value, not a buffer) first bin edge of the current RDF call
  size_t bytes;
};

#define MOUSE_REDUCE_BLOCKS 1024
#define MOUSE_WINDOW 32      // references of a home cell one work item of the P2 sweep takes (one item row per window)
#define MOUSE_SCAN_TILE 2048
#define MOUSE_MAX_BINS 8192

static inline size_t mouse_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves `ws` (may be NULL to only compute the size).  Identical on the query and on the use.
static inline Workspace mouse_carve(void *ws, int64_t n, int32_t ncell) {
  Workspace w;
  size_t off = 0;
  char *base = (char *)ws;
  size_t nn = (size_t)(n > 0 ? n : 1);
  size_t nc = (size_t)ncell + 1;
  size_t nsb = (nc + MOUSE_SCAN_TILE - 1) / MOUSE_SCAN_TILE + 1;
#define CARVE(field, type, count)                     \
  w.field = (type *)(base ? base + off : nullptr);    \
  off = mouse_align_up(off + sizeof(type) * (count), 256);
  CARVE(counters, int64_t, 32)     // counters, scan_blocks, cell_count and cell_bbox are adjacent: cleared by one memset
  CARVE(scan_blocks, int32_t, nsb)
  CARVE(cell_count, int32_t, nc)
  CARVE(cell_bbox, uint32_t, 6 * nc)
  CARVE(cell_start, int32_t, nc)
  CARVE(items, int32_t, 44 * (nc + nn / MOUSE_WINDOW + 1))
  CARVE(cell_of, int32_t, nn)
  CARVE(rank_of, int32_t, nn)
  CARVE(rec_tmp, Rec, nn)
  CARVE(idx_tmp, int32_t, nn)
  CARVE(rec_sorted, Rec, nn)
  CARVE(idx_sorted, int32_t, nn)
  CARVE(cell_sorted, int32_t, nn)
  CARVE(flag_sorted, uint8_t, nn)
  CARVE(acc_fix, long long, nn)
  CARVE(acc_cnt, int32_t, nn)
  CARVE(partials, double, 3 * MOUSE_REDUCE_BLOCKS)
  CARVE(edges, double, MOUSE_MAX_BINS + 1)
  CARVE(thr, double, MOUSE_MAX_BINS + 2)
  w.rdf_first = 0.0;
#undef CARVE
  w.bytes = off;
  return w;
}

// bytes from the start of the counters to the end of cell_bbox: what a cell build clears before it starts
static inline size_t mouse_clear_bytes(const Workspace &w, int32_t ncell) {
  return (size_t)((const char *)(w.cell_bbox + 6 * ((size_t)ncell + 1)) - (const char *)w.counters);
}

// order-preserving map float -> uint32 (a < b  <=>  key(a) < key(b)) and back
__host__ __device__ __forceinline__ uint32_t mouse_float_key(float x) {
#ifdef __CUDA_ARCH__
  const uint32_t b = __float_as_uint(x);
#else
  uint32_t b;
  memcpy(&b, &x, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float mouse_key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------ grid passed by value to kernels
struct GridDev {
  double box[3];
  double cw[3];
  double rc2;
  int nc[3];
  int ncell;
  float cwf[3];
  float boxf[3], inv_boxf[3];
  unsigned magic_x, magic_xy;   // floor(2^32 / nx), floor(2^32 / (nx * ny)): division-free cell index decoding
  float rc2_lo, rc2_hi;  // FP32 guard band: d2 < lo sure hit, d2 > hi sure miss
  // triclinic boxes (all zero for an orthorhombic grid): tilt factors xy, xz, yz, their FP32 copies, the ratios
  // xy/ly, xz/lz, yz/lz (fractional -> Cartesian) and the Cartesian centre of the cell H (1/2, 1/2, 1/2)
  double tilt[3];
  float tiltf[3];
  float tratio[3];
  double ctr_off[3];
  int tri;
};

static inline GridDev mouse_grid_dev(const mouse_grid_t *g) {
  GridDev d;
  for (int a = 0; a < 3; ++a) {
    d.box[a] = g->box[a];
    d.cw[a] = g->cell_w[a];
    d.nc[a] = g->ncell_axis[a];
    d.cwf[a] = (float)g->cell_w[a];
    d.boxf[a] = (float)g->box[a];
    d.inv_boxf[a] = (float)(1.0 / g->box[a]);
  }
  d.rc2 = g->rc2;
  d.ncell = g->ncell;
  d.magic_x = (unsigned)(4294967296ULL / (unsigned long long)(d.nc[0] > 1 ? d.nc[0] : 2));
  d.magic_xy = (unsigned)(4294967296ULL / (unsigned long long)(d.nc[0] * d.nc[1] > 1 ? d.nc[0] * d.nc[1] : 2));
  double lo = g->rc2 * (1.0 - (double)g->guard), hi = g->rc2 * (1.0 + (double)g->guard);
  d.rc2_lo = nextafterf((float)lo, 0.0f);
  d.rc2_hi = nextafterf((float)hi, 3.0e38f);
  d.tri = g->triclinic;
  for (int a = 0; a < 3; ++a) {
    d.tilt[a] = d.tri ? g->tilt[a] : 0.0;
    d.tiltf[a] = (float)d.tilt[a];
  }
  d.tratio[0] = (float)(d.tilt[0] / g->box[1]);
  d.tratio[1] = (float)(d.tilt[1] / g->box[2]);
  d.tratio[2] = (float)(d.tilt[2] / g->box[2]);
  d.ctr_off[0] = 0.5 * (g->box[0] + d.tilt[0] + d.tilt[1]);
  d.ctr_off[1] = 0.5 * (g->box[1] + d.tilt[2]);
  d.ctr_off[2] = 0.5 * g->box[2];
  return d;
}

// ------------------------------------------------------------------ exact (reference-order) arithmetic
// np.around(x) == rint (half to even); `dr - L * around(dr / L)` un-fused, true division
// (ordering_functions.py:98-100).
__device__ __forceinline__ double mouse_trim(double dr, double L) {
  return __dsub_rn(dr, __dmul_rn(L, rint(__ddiv_rn(dr, L))));
}

// rsq of ordering_functions.py:95-101 for centres (cix..) and (cjx..): (x^2 + y^2) + z^2
__device__ __forceinline__ double mouse_rsq_exact(double cix, double ciy, double ciz, double cjx,
                                                  double cjy, double cjz, const double *box) {
  double tx = mouse_trim(__dsub_rn(cjx, cix), box[0]);
  double ty = mouse_trim(__dsub_rn(cjy, ciy), box[1]);
  double tz = mouse_trim(__dsub_rn(cjz, ciz), box[2]);
  return __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
}

// The same for a triclinic box (build-defined): brick representative, axis by axis from z to x; tilt = {xy, xz, yz}.
// With tilt = 0 the three extra products vanish and the value is bit-identical to mouse_rsq_exact.
__device__ __forceinline__ void mouse_min_image_tri(double &dx, double &dy, double &dz, const double *box,
                                                    const double *tilt) {
  const double nz = rint(__ddiv_rn(dz, box[2]));
  dz = __dsub_rn(dz, __dmul_rn(box[2], nz));
  dy = __dsub_rn(dy, __dmul_rn(tilt[2], nz));
  dx = __dsub_rn(dx, __dmul_rn(tilt[1], nz));
  const double ny = rint(__ddiv_rn(dy, box[1]));
  dy = __dsub_rn(dy, __dmul_rn(box[1], ny));
  dx = __dsub_rn(dx, __dmul_rn(tilt[0], ny));
  const double nx = rint(__ddiv_rn(dx, box[0]));
  dx = __dsub_rn(dx, __dmul_rn(box[0], nx));
}
__device__ __forceinline__ double mouse_rsq_exact_g(double cix, double ciy, double ciz, double cjx, double cjy,
                                                    double cjz, const double *box, const double *tilt, int tri) {
  if (!tri) return mouse_rsq_exact(cix, ciy, ciz, cjx, cjy, cjz, box);
  double tx = __dsub_rn(cjx, cix), ty = __dsub_rn(cjy, ciy), tz = __dsub_rn(cjz, ciz);
  mouse_min_image_tri(tx, ty, tz, box, tilt);
  return __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
}

// un-fused dot product of ordering_functions.py:112: (bx*bx' + by*by') + bz*bz'
__device__ __forceinline__ double mouse_dot_exact(double ax, double ay, double az, double bx,
                                                  double by, double bz) {
  return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MOUSE_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MOUSE_FULL_MASK, v, o);
  return v;
}
