// Shared machinery of the half-shell sweeps (P2: p2_sweep.cu, RDF: rdf_sweep.cu): one warp per home cell, the
// candidates of the home cell and its 13 forward neighbour cells are contiguous ranges of the cell-sorted arrays,
// fetched with 1-D TMA bulk copies; the range table itself comes through a cp.async software pipeline.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------ PTX: mbarrier + 1-D TMA bulk copy
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MB_DONE;\n"
      "bra MB_WAIT;\n"
      "MB_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (TMA unit), completion counted in bytes on `bar`; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct HsItem {
  int home, hs, nhome, T;   // home cell, its first sorted slot, its bonds, candidates of the 14 cells
  int rsrc, rlen, rt;       // lane L < 10: range L of the candidate concatenation (first slot, length, offset t)
  float hx, hy, hz;         // home cell centre (frame of the stored FP32 coordinates)
  int win;                  // window of references this item covers: [win * MOUSE_WINDOW, ...) of the home cell (P2)
  int boundary;             // some neighbour cell is a periodic image
};

__device__ __forceinline__ void cp_async4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Range table of home cell tk.  Rows of the forward half shell: (dy,dz) = (0,0) cells x..x+1 (home first), (1,0),
// (-1,1), (0,1), (1,1) cells x-1..x+1; a row that crosses the periodic boundary in x splits into two ranges
// (part 0 / part 1).  Lane L < 10 owns range L = 2 * row + part: cells [lo, hi] of row base rb (empty when lo > hi).
// tk -> (cx, cy, cz) without integer division: multiply-high by floor(2^32 / d), then at most two corrections
__device__ __forceinline__ void hs_divmod(unsigned n, unsigned d, unsigned magic, int &q, int &r) {
  unsigned qq = __umulhi(n, magic);
  unsigned rr = n - qq * d;
  if (rr >= d) {
    rr -= d;
    qq += 1;
  }
  if (rr >= d) {
    rr -= d;
    qq += 1;
  }
  q = (int)qq;
  r = (int)rr;
}
__device__ __forceinline__ void hs_cell_coords(const GridDev &g, int tk, int &cx, int &cy, int &cz) {
  int rem;
  hs_divmod((unsigned)tk, (unsigned)(g.nc[0] * g.nc[1]), g.magic_xy, cz, rem);
  hs_divmod((unsigned)rem, (unsigned)g.nc[0], g.magic_x, cy, cx);
}

__device__ __forceinline__ void hs_range_cells(const GridDev &g, int cx, int cy, int cz, int lane, int &rb, int &lo,
                                               int &hi) {
  const int nx = g.nc[0], ny = g.nc[1], nz = g.nc[2];
  const int row = lane >> 1, part = lane & 1;
  const int dy = (row == 1 || row == 4) ? 1 : (row == 2 ? -1 : 0);
  const int dz = row >= 2 ? 1 : 0;
  int yc = cy + dy, zc = cz + dz;
  yc += yc < 0 ? ny : 0;
  yc -= yc >= ny ? ny : 0;
  zc -= zc >= nz ? nz : 0;
  rb = (zc * ny + yc) * nx;
  const int xa = row == 0 ? cx : cx - 1, xb = cx + 1;
  if (xa < 0) {
    lo = part ? 0 : nx - 1;
    hi = part ? xb : nx - 1;
  } else if (xb >= nx) {
    lo = part ? 0 : xa;
    hi = part ? xb - nx : nx - 1;
  } else {
    lo = part ? 1 : xa;
    hi = part ? 0 : xb;
  }
}

// Item rows (44 ints each, written by cell_items_kernel - one thread per cell - in the order the scan assigned:
// heaviest cells first; read by the sweeps through a ticket counter):
//   [0..9] first sorted slot of range L, [10..19] its length, [20..29] its offset t in the candidate concatenation,
//   [30] bonds / atoms in the home cell | boundary << 30 (some neighbour cell is a periodic image),
//   [32..34] / [35..37] lower / upper corner of the bounding box of the cell's own items (FP32 bit patterns, frame of
//   the stored coordinates: wrapped - L/2), [38], [39], [31] the cell centre x, y, z in the same frame,
//   [40] window index of the row (P2: a cell with more than MOUSE_WINDOW items has one row per window), [41..43] unused.
#define HS_ITEM_INTS 44

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}

// Periodic image of (px, py, pz) nearest to the home cell centre h, in the brick sense (axis by axis from z to x; a
// shift along an axis also removes the tilt components of the lattice vector from the lower axes).  FP32, un-fused.
__device__ __forceinline__ void hs_nearest_image(float &px, float &py, float &pz, float hx, float hy, float hz,
                                                 const GridDev &g) {
  const float nz = rintf((pz - hz) * g.inv_boxf[2]);
  pz -= g.boxf[2] * nz;
  py -= g.tiltf[2] * nz;
  px -= g.tiltf[1] * nz;
  const float ny = rintf((py - hy) * g.inv_boxf[1]);
  py -= g.boxf[1] * ny;
  px -= g.tiltf[0] * ny;
  const float nx = rintf((px - hx) * g.inv_boxf[0]);
  px -= g.boxf[0] * nx;
}

// fetch stage 2: start the (asynchronous) copy of home cell tk's item row into s_raw
__device__ __forceinline__ void hs_raw_issue(const int32_t *__restrict__ items, int lane, int tk, int *s_raw) {
  if (lane < HS_ITEM_INTS / 4) cp_async16(s_raw + 4 * lane, items + (size_t)tk * HS_ITEM_INTS + 4 * lane);
  cp_async_commit();
}

// fetch stage 3: the item row of home cell tk has landed in s_raw; false when the cell is empty
__device__ __forceinline__ bool hs_finalize(const GridDev &g, int lane, int tk, const int *s_raw, HsItem &it) {
  cp_async_wait_all();
  __syncwarp();
  const int l = lane < 10 ? lane : 0;
  const int src = s_raw[l], len = lane < 10 ? s_raw[10 + l] : 0, rt = s_raw[20 + l];
  const int T = s_raw[29] + s_raw[19];
  const int size = s_raw[30];
  const float hx = __int_as_float(s_raw[38]), hy = __int_as_float(s_raw[39]), hz = __int_as_float(s_raw[31]);
  const int win = s_raw[40];
  __syncwarp();
  const int nhome = size & 0x3fffffff;
  it.home = tk;
  it.hs = __shfl_sync(MOUSE_FULL_MASK, src, 0);      // range 0 starts with the home cell
  it.nhome = nhome;
  it.T = T;
  it.rsrc = src;
  it.rlen = len;
  it.rt = rt;
  it.hx = hx;          // cell centre, Cartesian, frame of the stored coordinates (written by the canon kernel)
  it.hy = hy;
  it.hz = hz;
  it.boundary = (size >> 30) & 1;
  it.win = win;
  return nhome > 0;
}
