// C-ABI entry points of libmouse_b200 (see include/mouse_b200.h for the contract).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "common.cuh"

// launchers implemented next to their kernels
int mouse_launch_bond_build(const double *pos, const int32_t *bond_atoms, int64_t n, const double *box,
                            const double *tilt, double *vec, double *ctr, cudaStream_t st);
int mouse_launch_cell_build(int mode, const double *xyz, const double *vec, const int32_t *code,
                            const int32_t *code2, const uint8_t *nbr_mask, const uint8_t *ref_mask, int64_t n,
                            const mouse_grid_t *grid, const Workspace &w, cudaStream_t st);
int mouse_launch_frame_build(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                             const uint8_t *ref_mask, int64_t n, const mouse_grid_t *grid, const Workspace &w,
                             double *vec, double *ctr, double *out_sum, int32_t *out_cnt, double *out_mean,
                             int init_listed, cudaStream_t st);
int mouse_launch_p2_sweep(const double *vec, const double *ctr, int64_t n, const mouse_grid_t *grid, uint32_t flags,
                          const Workspace &w, double *out_sum, int32_t *out_cnt, const int64_t *row_ptr,
                          int32_t *nbr_idx, int64_t list_cap, int frame_mode, int *used_fast, cudaStream_t st);
int mouse_launch_p2_finish_reduce(int64_t n, const mouse_grid_t *grid, const Workspace &w, double *out_sum,
                                  int32_t *out_cnt, double *out_mean, mouse_p2_totals_t *totals, cudaStream_t st);
int mouse_launch_row_ptr(const int32_t *cnt, int64_t n, int64_t *row_ptr, cudaStream_t st);
int mouse_launch_p2_reduce(const double *sum, const int32_t *cnt, int64_t n, const Workspace &w, double *out_mean,
                           mouse_p2_totals_t *totals, int64_t *hist, double smin, double smax, int32_t nbins,
                           cudaStream_t st);
int mouse_launch_rdf_sweep(const mouse_grid_t *grid, uint32_t flags, int64_t n, int32_t ntypes, const double *edges_dev,
                           int32_t nbin, const Workspace &w, int64_t *hist, cudaStream_t st);
int mouse_launch_director_p2(const double *vec, const uint8_t *mask, int64_t n, const double *director, double *out_s,
                             int64_t *hist, double smin, double smax, int32_t nbins, cudaStream_t st);
int mouse_launch_atom_colour(const double *mean, const int32_t *cnt, const int64_t *row_ptr, const int32_t *bond_of,
                             int64_t n_atoms, int32_t direct, int32_t *code, cudaStream_t st);
int mouse_p2_uses_fast(const mouse_grid_t *grid, int64_t n, uint32_t flags, bool lists);

static thread_local char g_err[512] = "";

void mouse_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char *mouse_version(void) { return "mouse_b200 0.1 (sm_100a)"; }
const char *mouse_last_error(void) { return g_err; }

static int grid_plan_impl(const char *who, const double *box, const double *tilt, double rcut, double rc2,
                          int64_t n_items, mouse_grid_t *grid) {
  if (!box || !grid) {
    mouse_set_error("%s: NULL argument", who);
    return MOUSE_EINVAL;
  }
  memset(grid, 0, sizeof(*grid));
  for (int a = 0; a < 3; ++a) {
    if (!(box[a] > 0.0)) {
      mouse_set_error("%s: box[%d]=%g must be positive", who, a, box[a]);
      return MOUSE_EINVAL;
    }
  }
  // distance between opposite faces of the box along every lattice direction (= the edge for an orthorhombic box)
  double wid[3] = {box[0], box[1], box[2]};
  double xy = 0.0, xz = 0.0, yz = 0.0;
  if (tilt) {
    xy = tilt[0], xz = tilt[1], yz = tilt[2];
    const double bcx = box[1] * box[2], bcy = -xy * box[2], bcz = xy * yz - box[1] * xz;   // b x c
    wid[0] = box[0] * box[1] * box[2] / sqrt(bcx * bcx + bcy * bcy + bcz * bcz);
    wid[1] = box[1] * box[2] / sqrt(box[2] * box[2] + yz * yz);                            // V / |a x c|
  }
  double cells = 1.0;
  int nax[3];
  for (int a = 0; a < 3; ++a) {
    int k = 1;
    if (rcut > 0.0) {
      double q = wid[a] / (rcut * (1.0 + 1e-6));
      k = q >= 1.0 ? (q > 1024.0 ? 1024 : (int)q) : 1;
    }
    nax[a] = k;
    cells *= k;
  }
  // dilute systems: do not allocate many more cells than items
  double cap = 4.0 * (double)(n_items > 16 ? n_items : 16);
  if (cap > 2.0e8) cap = 2.0e8;
  if (cells > cap) {
    double f = cbrt(cap / cells);
    for (int a = 0; a < 3; ++a) {
      int k = (int)(nax[a] * f);
      nax[a] = k < 1 ? 1 : k;
    }
  }
  // sparse cells: the half-shell sweeps pay a fixed cost per home cell (fetch, TMA, cull, register load, flush), which
  // is amortised best at ~10 items per cell; any cell at least rcut wide is a valid cell list.
  // Small systems keep the fine grid: they need the home cells for parallelism, not for amortisation.
  {
    const double target = 9.5;
    double now = (double)nax[0] * nax[1] * nax[2];
    if (rcut > 0.0 && n_items >= 65536 && now * target > (double)n_items) {
      double f = cbrt((double)n_items / target / now);
      for (int a = 0; a < 3; ++a) {
        int k = (int)(nax[a] * f);
        int lo = nax[a] < 3 ? nax[a] : 3;
        nax[a] = k < lo ? lo : k;
      }
    }
  }
  int fast = rcut > 0.0;
  for (int a = 0; a < 3; ++a) {
    grid->box[a] = box[a];
    grid->ncell_axis[a] = nax[a];
    grid->cell_w[a] = box[a] / nax[a];
    if (nax[a] < 3) fast = 0;
  }
  grid->ncell = nax[0] * nax[1] * nax[2];
  grid->rcut = rcut;
  grid->rc2 = rc2;
  grid->tilt[0] = xy, grid->tilt[1] = xz, grid->tilt[2] = yz;
  grid->triclinic = tilt ? 1 : 0;
  // FP32 error budget of the fast sweep (DESIGN.md "guard band"): coordinates are stored as FP32 of
  // (wrapped - centre), one rounded periodic shift, one subtraction: |delta dx| <= 2.5 * 2^-24 * L, so the
  // relative error of d2 at the cutoff is <= 5.2e-7 * (Lmax / rcut) + 1.8e-7; keep a 4x margin, never below 2^-15.
  // A triclinic shift is up to three rounded steps per coordinate on a Cartesian extent lx + |xy| + |xz|: 3x.
  double ext[3] = {box[0] + fabs(xy) + fabs(xz), box[1] + fabs(yz), box[2]};
  double lmax = ext[0] > ext[1] ? (ext[0] > ext[2] ? ext[0] : ext[2]) : (ext[1] > ext[2] ? ext[1] : ext[2]);
  double guard = rcut > 0.0 ? 4.0 * ((tilt ? 3.0 : 1.0) * 5.2e-7 * (lmax / rcut) + 2.0e-7) : 1.0;
  if (guard < 3.0517578125e-05) guard = 3.0517578125e-05;
  if (guard > 1.0e-2) fast = 0;
  if (tilt && fast) {
    // the sweep moves a candidate to the image nearest to the home cell's centre with the brick rule: that is the
    // image a reference of the home cell needs only if half a cell plus the cutoff fits half the box edge on every
    // Cartesian axis (for an orthorhombic box this is the ">= 3 cells" rule)
    const double cext[3] = {box[0] / nax[0] + fabs(xy) / nax[1] + fabs(xz) / nax[2], box[1] / nax[1] + fabs(yz) / nax[2],
                            box[2] / nax[2]};
    for (int a = 0; a < 3; ++a)
      if (0.5 * cext[a] + rcut * (1.0 + 2.0 * guard) > 0.5 * box[a] * (1.0 - 1e-9)) fast = 0;
  }
  grid->guard = (float)guard;
  grid->fast = fast;
  return MOUSE_OK;
}

int mouse_grid_plan(const double *box, double rcut, double rc2, int64_t n_items, mouse_grid_t *grid) {
  return grid_plan_impl("mouse_grid_plan", box, nullptr, rcut, rc2, n_items, grid);
}

int mouse_grid_plan_triclinic(const double *box, const double *tilt, double rcut, double rc2, int64_t n_items,
                              mouse_grid_t *grid) {
  if (!tilt) {
    mouse_set_error("mouse_grid_plan_triclinic: NULL tilt");
    return MOUSE_EINVAL;
  }
  return grid_plan_impl("mouse_grid_plan_triclinic", box, tilt, rcut, rc2, n_items, grid);
}

size_t mouse_workspace_bytes(int64_t n_items, const mouse_grid_t *grid) {
  if (!grid || n_items < 0) return 0;
  return mouse_carve(nullptr, n_items, grid->ncell).bytes;
}

static int check_ws(const char *who, int64_t n, const mouse_grid_t *grid, void *ws, size_t ws_bytes, Workspace *w) {
  if (!grid || !ws) {
    mouse_set_error("%s: NULL grid/workspace", who);
    return MOUSE_EINVAL;
  }
  if (n < 0 || n >= (1LL << 30)) {
    mouse_set_error("%s: n=%lld out of range", who, (long long)n);
    return MOUSE_EINVAL;
  }
  *w = mouse_carve(ws, n, grid->ncell);
  if (ws_bytes < w->bytes) {
    mouse_set_error("%s: workspace %zu < required %zu bytes", who, ws_bytes, w->bytes);
    return MOUSE_EWORKSPACE;
  }
  if (((uintptr_t)ws & 255u) != 0) {
    mouse_set_error("%s: workspace must be 256-byte aligned", who);
    return MOUSE_EINVAL;
  }
  return MOUSE_OK;
}

int mouse_bond_build(const double *pos, const int32_t *bond_atoms, int64_t n_bonds, const double *box, double *vec,
                     double *ctr, void *stream) {
  if (!pos || !bond_atoms || !box || !vec || !ctr) {
    if (n_bonds == 0) return MOUSE_OK;
    mouse_set_error("mouse_bond_build: NULL argument");
    return MOUSE_EINVAL;
  }
  return mouse_launch_bond_build(pos, bond_atoms, n_bonds, box, nullptr, vec, ctr, (cudaStream_t)stream);
}

int mouse_bond_build_triclinic(const double *pos, const int32_t *bond_atoms, int64_t n_bonds, const double *box,
                               const double *tilt, double *vec, double *ctr, void *stream) {
  if (!pos || !bond_atoms || !box || !tilt || !vec || !ctr) {
    if (n_bonds == 0) return MOUSE_OK;
    mouse_set_error("mouse_bond_build_triclinic: NULL argument");
    return MOUSE_EINVAL;
  }
  return mouse_launch_bond_build(pos, bond_atoms, n_bonds, box, tilt, vec, ctr, (cudaStream_t)stream);
}

int mouse_cell_build(const double *vec, const double *ctr, const int32_t *mol, const uint8_t *nbr_mask,
                     const uint8_t *ref_mask, int64_t n_bonds, const mouse_grid_t *grid, void *ws, size_t ws_bytes,
                     void *stream) {
  Workspace w;
  int rc = check_ws("mouse_cell_build", n_bonds, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  return mouse_launch_cell_build(0, ctr, vec, mol, nullptr, nbr_mask, ref_mask, n_bonds, grid, w,
                                 (cudaStream_t)stream);
}

int mouse_p2_sweep(const double *vec, const double *ctr, int64_t n_bonds, const mouse_grid_t *grid, uint32_t flags,
                   void *ws, size_t ws_bytes, double *out_sum, int32_t *out_cnt, void *stream) {
  if (!out_sum || !out_cnt) {
    mouse_set_error("mouse_p2_sweep: NULL argument");
    return MOUSE_EINVAL;
  }
  Workspace w;
  int rc = check_ws("mouse_p2_sweep", n_bonds, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  return mouse_launch_p2_sweep(vec, ctr, n_bonds, grid, flags, w, out_sum, out_cnt, nullptr, nullptr, 0, 0, nullptr,
                               (cudaStream_t)stream);
}

int mouse_p2_pair_lists(const double *vec, const double *ctr, int64_t n_bonds, const mouse_grid_t *grid,
                        uint32_t flags, void *ws, size_t ws_bytes, const int32_t *cnt, int64_t *row_ptr,
                        int32_t *nbr_idx, int64_t capacity, void *stream) {
  if (!cnt || !row_ptr || !nbr_idx || capacity < 0) {
    mouse_set_error("mouse_p2_pair_lists: NULL argument / negative capacity");
    return MOUSE_EINVAL;
  }
  Workspace w;
  int rc = check_ws("mouse_p2_pair_lists", n_bonds, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  if ((rc = mouse_launch_row_ptr(cnt, n_bonds, row_ptr, (cudaStream_t)stream))) return rc;
  // rank_of is free after the cell build: scratch "entries written so far" per row.  Entries at or beyond
  // `capacity` are not written (the kernel checks every store against it)
  return mouse_launch_p2_sweep(vec, ctr, n_bonds, grid, flags, w, nullptr, w.rank_of, row_ptr, nbr_idx, capacity, 0,
                               nullptr, (cudaStream_t)stream);
}

int mouse_p2_reduce(const double *sum, const int32_t *cnt, int64_t n_bonds, const mouse_grid_t *grid, void *ws,
                    size_t ws_bytes, double *out_mean, mouse_p2_totals_t *totals, int64_t *hist, double smin,
                    double smax, int32_t nbins, void *stream) {
  if (!sum || !cnt || !out_mean || !totals) {
    mouse_set_error("mouse_p2_reduce: NULL argument");
    return MOUSE_EINVAL;
  }
  Workspace w;
  int rc = check_ws("mouse_p2_reduce", n_bonds, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  return mouse_launch_p2_reduce(sum, cnt, n_bonds, w, out_mean, totals, hist, smin, smax, nbins,
                                (cudaStream_t)stream);
}

int mouse_p2_frame_timed(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                         const uint8_t *ref_mask, int64_t n_bonds, const mouse_grid_t *grid, uint32_t flags, void *ws,
                         size_t ws_bytes, double *vec, double *ctr, double *sum, int32_t *cnt, double *mean,
                         mouse_p2_totals_t *totals, void *ev_sweep_begin, void *ev_sweep_end, void *stream) {
  Workspace w;
  int rc = check_ws("mouse_p2_frame", n_bonds, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (!pos || !bond_atoms || !vec || !ctr || !sum || !cnt || !mean || !totals) {
    if (n_bonds > 0) {
      mouse_set_error("mouse_p2_frame: NULL argument");
      return MOUSE_EINVAL;
    }
  }
  // K1 + K2 fused (bond vectors, cell list, outputs initialised) -> K3 -> K4 fused with the accumulator conversion
  // (the fast sweep's K4 writes every listed bond itself; the all-FP64 path needs all outputs initialised)
  const int fast_path = mouse_p2_uses_fast(grid, n_bonds, flags, false);
  if ((rc = mouse_launch_frame_build(pos, bond_atoms, mol, nbr_mask, ref_mask, n_bonds, grid, w, vec, ctr, sum, cnt,
                                     mean, !fast_path, st)))
    return rc;
  int used_fast = 0;
  if (ev_sweep_begin) MOUSE_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev_sweep_begin, st));
  if ((rc = mouse_launch_p2_sweep(vec, ctr, n_bonds, grid, flags, w, sum, cnt, nullptr, nullptr, 0, 1, &used_fast, st)))
    return rc;
  if (ev_sweep_end) MOUSE_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev_sweep_end, st));
  if (used_fast) return mouse_launch_p2_finish_reduce(n_bonds, grid, w, sum, cnt, mean, totals, st);
  return mouse_launch_p2_reduce(sum, cnt, n_bonds, w, mean, totals, nullptr, 0.0, 0.0, 0, st);
}

int mouse_p2_frame(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                   const uint8_t *ref_mask, int64_t n_bonds, const mouse_grid_t *grid, uint32_t flags, void *ws,
                   size_t ws_bytes, double *vec, double *ctr, double *sum, int32_t *cnt, double *mean,
                   mouse_p2_totals_t *totals, void *stream) {
  return mouse_p2_frame_timed(pos, bond_atoms, mol, nbr_mask, ref_mask, n_bonds, grid, flags, ws, ws_bytes, vec, ctr, sum,
                              cnt, mean, totals, nullptr, nullptr, stream);
}

// ---- a frame captured once into a CUDA graph: one launch per frame instead of a train of eight (the buffers, the
// grid and the flags are baked in; new positions are copied into `pos` by the caller before every launch)
struct FrameGraph {
  cudaGraphExec_t exec;
  int nodes;
};

int mouse_p2_frame_graph_create(const double *pos, const int32_t *bond_atoms, const int32_t *mol,
                                const uint8_t *nbr_mask, const uint8_t *ref_mask, int64_t n_bonds,
                                const mouse_grid_t *grid, uint32_t flags, void *ws, size_t ws_bytes, double *vec,
                                double *ctr, double *sum, int32_t *cnt, double *mean, mouse_p2_totals_t *totals,
                                void **out_graph) {
  if (!out_graph) {
    mouse_set_error("mouse_p2_frame_graph_create: NULL out_graph");
    return MOUSE_EINVAL;
  }
  *out_graph = nullptr;
  cudaStream_t st = nullptr;
  MOUSE_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  int rc = MOUSE_OK;
  if (e == cudaSuccess) {
    rc = mouse_p2_frame(pos, bond_atoms, mol, nbr_mask, ref_mask, n_bonds, grid, flags, ws, ws_bytes, vec, ctr, sum, cnt,
                        mean, totals, st);
    e = cudaStreamEndCapture(st, &g);
  }
  cudaGraphExec_t exec = nullptr;
  size_t nodes = 0;
  if (e == cudaSuccess && rc == MOUSE_OK) {
    cudaGraphGetNodes(g, nullptr, &nodes);
    e = cudaGraphInstantiate(&exec, g, 0);
  }
  if (g) cudaGraphDestroy(g);
  cudaStreamDestroy(st);
  if (rc != MOUSE_OK) return rc;
  if (e != cudaSuccess) {
    mouse_set_error("mouse_p2_frame_graph_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return MOUSE_ECUDA;
  }
  FrameGraph *fg = (FrameGraph *)malloc(sizeof(FrameGraph));
  fg->exec = exec;
  fg->nodes = (int)nodes;
  *out_graph = fg;
  return MOUSE_OK;
}

int mouse_p2_frame_graph_launch(void *graph, void *stream) {
  if (!graph) {
    mouse_set_error("mouse_p2_frame_graph_launch: NULL graph");
    return MOUSE_EINVAL;
  }
  MOUSE_CUDA_CHECK(cudaGraphLaunch(((FrameGraph *)graph)->exec, (cudaStream_t)stream));
  return MOUSE_OK;
}

int mouse_p2_frame_graph_nodes(void *graph) { return graph ? ((FrameGraph *)graph)->nodes : 0; }

int mouse_p2_frame_graph_destroy(void *graph) {
  if (!graph) return MOUSE_OK;
  cudaGraphExecDestroy(((FrameGraph *)graph)->exec);
  free(graph);
  return MOUSE_OK;
}

int mouse_rdf_frame_timed(const double *pos, const int32_t *type, const int32_t *mol, int64_t n_atoms, int32_t ntypes,
                          const mouse_grid_t *grid, uint32_t flags, const double *edges, int32_t nbin, void *ws,
                          size_t ws_bytes, int64_t *hist, void *ev_sweep_begin, void *ev_sweep_end, void *stream) {
  Workspace w;
  int rc = check_ws("mouse_rdf_frame", n_atoms, grid, ws, ws_bytes, &w);
  if (rc) return rc;
  if (!pos || !type || !edges || !hist || ntypes < 1 || nbin < 1 || nbin > MOUSE_MAX_BINS) {
    mouse_set_error("mouse_rdf_frame: bad argument (ntypes=%d nbin=%d)", ntypes, nbin);
    return MOUSE_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MOUSE_CUDA_CHECK(cudaMemcpyAsync(w.edges, edges, sizeof(double) * (size_t)(nbin + 1), cudaMemcpyHostToDevice, st));
  if ((rc = mouse_launch_cell_build(1, pos, nullptr, type, mol, nullptr, nullptr, n_atoms, grid, w, st))) return rc;
  w.rdf_first = edges[0];
  // the fast sweep's cutoff is the last edge: it needs a grid planned for exactly that cutoff
  if (!(grid->rcut == edges[nbin]) || !(edges[nbin] > edges[0])) flags |= MOUSE_FORCE_EXACT;
  if (ev_sweep_begin) MOUSE_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev_sweep_begin, st));
  rc = mouse_launch_rdf_sweep(grid, flags, n_atoms, ntypes, w.edges, nbin, w, hist, st);
  if (ev_sweep_end) MOUSE_CUDA_CHECK(cudaEventRecord((cudaEvent_t)ev_sweep_end, st));
  return rc;
}

int mouse_rdf_frame(const double *pos, const int32_t *type, const int32_t *mol, int64_t n_atoms, int32_t ntypes,
                    const mouse_grid_t *grid, uint32_t flags, const double *edges, int32_t nbin, void *ws,
                    size_t ws_bytes, int64_t *hist, void *stream) {
  return mouse_rdf_frame_timed(pos, type, mol, n_atoms, ntypes, grid, flags, edges, nbin, ws, ws_bytes, hist, nullptr,
                               nullptr, stream);
}

int mouse_director_p2(const double *vec, const uint8_t *mask, int64_t n_bonds, const double *director, double *out_s,
                      int64_t *hist, double smin, double smax, int32_t nbins, void *stream) {
  if (!vec || !director || !out_s || !hist || nbins < 1 || nbins > MOUSE_MAX_BINS) {
    if (n_bonds == 0 && hist && nbins >= 1 && nbins <= MOUSE_MAX_BINS)
      return mouse_launch_director_p2(vec, mask, 0, director, out_s, hist, smin, smax, nbins, (cudaStream_t)stream);
    mouse_set_error("mouse_director_p2: bad argument (nbins=%d)", nbins);
    return MOUSE_EINVAL;
  }
  return mouse_launch_director_p2(vec, mask, n_bonds, director, out_s, hist, smin, smax, nbins, (cudaStream_t)stream);
}

int mouse_atom_colour(const double *mean, const int32_t *cnt, const int64_t *row_ptr, const int32_t *bond_of,
                      int64_t n_atoms, int32_t direct, int32_t *code, void *stream) {
  if (n_atoms < 0 || (n_atoms > 0 && (!mean || !cnt || !row_ptr || !bond_of || !code))) {
    mouse_set_error("mouse_atom_colour: bad argument");
    return MOUSE_EINVAL;
  }
  return mouse_launch_atom_colour(mean, cnt, row_ptr, bond_of, n_atoms, direct, code, (cudaStream_t)stream);
}

// Debug / tuning: copies the 8 workspace counters to host (synchronises the stream).
int mouse_debug_counters(const mouse_grid_t *grid, int64_t n_items, const void *ws, int64_t *out8, void *stream) {
  Workspace w = mouse_carve((void *)ws, n_items, grid->ncell);
  MOUSE_CUDA_CHECK(cudaMemcpyAsync(out8, w.counters, sizeof(int64_t) * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  MOUSE_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return MOUSE_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ host-side ingestion helpers
static inline bool mouse_is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// One decimal number starting at s (no leading whitespace), token ends at whitespace or `end`.  Correctly rounded:
// numbers with <= 15 significant digits and a decimal exponent within +-22 take the exact fast path (integer
// mantissa times / divided by an exactly representable power of ten: one rounding), all others go through strtod.
// Returns the end of the token, or NULL when the token is not a number.
static inline const char *mouse_parse_token(const char *s, const char *end, double *value) {
  static const double p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char *tok = s;
  bool neg = false;
  if (s < end && (*s == '-' || *s == '+')) neg = (*s++ == '-');
  uint64_t m = 0;
  int digits = 0, frac = 0;
  bool any = false, lead = true;
  while (s < end && *s >= '0' && *s <= '9') {
    any = true;
    if (!(lead && *s == '0')) {
      lead = false;
      if (digits < 19) m = m * 10 + (uint64_t)(*s - '0');
      ++digits;
    }
    ++s;
  }
  if (s < end && *s == '.') {
    ++s;
    while (s < end && *s >= '0' && *s <= '9') {
      any = true;
      if (!(lead && *s == '0')) {
        lead = false;
        if (digits < 19) m = m * 10 + (uint64_t)(*s - '0');
        ++digits;
      }
      ++frac;
      ++s;
    }
  }
  int e10 = 0;
  bool fast = any;
  if (any && s < end && (*s == 'e' || *s == 'E')) {
    const char *save = s++;
    bool eneg = false;
    if (s < end && (*s == '-' || *s == '+')) eneg = (*s++ == '-');
    if (s < end && *s >= '0' && *s <= '9') {
      int e = 0;
      while (s < end && *s >= '0' && *s <= '9') {
        if (e < 10000) e = e * 10 + (*s - '0');
        ++s;
      }
      e10 = eneg ? -e : e;
    } else {
      s = save;      // "1e" / "1ex": not an exponent
      fast = false;
    }
  }
  const bool token_ends = s >= end || mouse_is_ws(*s);
  const int e = e10 - frac;
  if (fast && token_ends && digits <= 15 && e >= -22 && e <= 22) {
    double v = (double)m;                       // exact: m < 10^15 < 2^53
    v = e >= 0 ? v * p10[e] : v / p10[-e];      // one correctly rounded operation with an exact power of ten
    *value = neg ? -v : v;
    return s;
  }
  // everything else (long mantissas, large exponents, nan / inf, malformed tokens): the C library
  const char *t = tok;
  while (t < end && !mouse_is_ws(*t)) ++t;
  char buf[128];
  const size_t tl = (size_t)(t - tok);
  if (tl == 0 || tl >= sizeof(buf)) return nullptr;
  memcpy(buf, tok, tl);
  buf[tl] = 0;
  char *stop = nullptr;
  const double v = strtod(buf, &stop);
  if (stop != buf + tl) return nullptr;
  *value = v;
  return t;
}

extern "C" int64_t mouse_parse_doubles(const char *text, int64_t len, double *out, int64_t cap) {
  const char *s = text, *end = text + len;
  int64_t n = 0;
  for (;;) {
    while (s < end && mouse_is_ws(*s)) ++s;
    if (s >= end) break;
    if (n >= cap) return -((int64_t)(s - text) + 1);
    double v;
    const char *next = mouse_parse_token(s, end, &v);
    if (!next) return -((int64_t)(s - text) + 1);
    out[n++] = v;
    s = next;
  }
  return n;
}

// The same over several host threads: the text is cut at whitespace into one chunk per thread, a first pass counts the
// tokens of every chunk (so that every chunk knows where its values go), a second pass parses the chunks in place.
// Token boundaries do not depend on the cut, every token is converted by the same code as above: the values (and the
// error position) are those of the single-threaded call.
extern "C" int64_t mouse_parse_doubles_mt(const char *text, int64_t len, double *out, int64_t cap, int32_t nthreads) {
  if (nthreads > 64) nthreads = 64;
  if (nthreads <= 1 || len < (int64_t)1 << 16) return mouse_parse_doubles(text, len, out, cap);
  std::vector<int64_t> cut(nthreads + 1), count(nthreads, 0), result(nthreads, 0);
  cut[0] = 0;
  cut[nthreads] = len;
  for (int t = 1; t < nthreads; ++t) {
    int64_t c = len * t / nthreads;
    if (c < cut[t - 1]) c = cut[t - 1];
    while (c < len && !mouse_is_ws(text[c])) ++c;     // never inside a token
    cut[t] = c;
  }
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t]() {
        int64_t n = 0;
        bool in_tok = false;
        for (int64_t k = cut[t]; k < cut[t + 1]; ++k) {
          const bool w = mouse_is_ws(text[k]);
          n += (!w && !in_tok);
          in_tok = !w;
        }
        count[t] = n;
      });
    for (auto &th : pool) th.join();
  }
  std::vector<int64_t> first(nthreads + 1, 0);
  for (int t = 0; t < nthreads; ++t) first[t + 1] = first[t] + count[t];
  if (first[nthreads] > cap) {
    // the single-threaded call reports the first token that does not fit
    return mouse_parse_doubles(text, len, out, cap);
  }
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
      pool.emplace_back([&, t]() {
        result[t] = mouse_parse_doubles(text + cut[t], cut[t + 1] - cut[t], out + first[t], count[t]);
      });
    for (auto &th : pool) th.join();
  }
  for (int t = 0; t < nthreads; ++t)
    if (result[t] < 0) return -((-result[t] - 1) + cut[t] + 1);     // offset of the bad token in the whole text
  return first[nthreads];
}

// A section of a LAMMPS data file (Atoms, Bonds: lammpack_types.py:400-464 reads them line by line with
// AtomFromLmpDataLine / BondFromLmpDataLine, lammpack_lmp_functions.py:3-33): `nrows` consecutive lines of exactly
// `ncol` whitespace-separated tokens; the `nwant` columns cols[] go to the arrays outs[] - kind 0: a decimal number as a
// correctly rounded double (mouse_parse_token), kind 1: an integer in CANONICAL form ([-]digits, no plus sign, no
// leading zeros, at most 18 digits) as int64, so that printing the value gives the token back (types and molecule ids
// are strings in the reference).  Tokens of the other columns are skipped unread.  *consumed = bytes up to and including
// the newline of the last line.  Returns nrows, or a negative number when the block is not of that shape (blank or
// ragged line, a wanted token of another form, fewer lines): the caller then takes the general line-by-line path.
extern "C" int64_t mouse_parse_table(const char *text, int64_t len, int64_t nrows, int32_t ncol, int32_t nwant,
                                     const int32_t *cols, const int32_t *kinds, void *const *outs, int64_t *consumed,
                                     int32_t nthreads) {
  if (!text || len < 0 || nrows < 0 || ncol < 1 || nwant < 0 || nwant > 16 || (nwant && (!cols || !kinds || !outs)))
    return -1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if (nrows < 4096) nthreads = 1;
  std::vector<int> want(ncol, -1);
  for (int w = 0; w < nwant; ++w) {
    if (cols[w] < 0 || cols[w] >= ncol || want[cols[w]] >= 0 || (kinds[w] != 0 && kinds[w] != 1)) return -1;
    want[cols[w]] = w;
  }
  // extent of the block and the first byte of every thread's share of the lines
  std::vector<int64_t> cut(nthreads + 1, 0), row0(nthreads + 1, 0);
  {
    const char *p = text, *end = text + len;
    int t = 1;
    for (int64_t r = 0; r < nrows; ++r) {
      if (p >= end) return -1 - len;
      while (t < nthreads && r == nrows * t / nthreads) {
        cut[t] = p - text;
        row0[t] = r;
        ++t;
      }
      const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
      p = nl ? nl + 1 : end;
    }
    for (; t <= nthreads; ++t) {
      cut[t] = p - text;
      row0[t] = nrows;
    }
  }
  if (consumed) *consumed = cut[nthreads];
  std::vector<int64_t> result(nthreads, 0);
  auto parse_chunk = [&](int t) {
    const char *s = text + cut[t], *end = text + cut[t + 1];
    for (int64_t row = row0[t]; row < row0[t + 1]; ++row) {
      const char *eol = (const char *)memchr(s, '\n', (size_t)(end - s));
      if (!eol) eol = end;
      const char *q = s;
      int col = 0;
      for (;;) {
        while (q < eol && mouse_is_ws(*q)) ++q;
        if (q >= eol) break;
        if (col >= ncol) {
          result[t] = -((int64_t)(s - text) + 2);
          return;
        }
        const int w = want[col];
        if (w < 0) {
          while (q < eol && !mouse_is_ws(*q)) ++q;
        } else if (kinds[w] == 0) {
          double v;
          const char *next = mouse_parse_token(q, eol, &v);
          if (!next) {
            result[t] = -((int64_t)(s - text) + 2);
            return;
          }
          ((double *)outs[w])[row] = v;
          q = next;
        } else {
          const char *d = q;
          const bool neg = *d == '-';
          if (neg) ++d;
          const char *first = d;
          int64_t v = 0;
          while (d < eol && *d >= '0' && *d <= '9') v = v * 10 + (*d++ - '0');
          const int nd = (int)(d - first);
          if (nd < 1 || nd > 18 || (nd > 1 && *first == '0') || (neg && nd == 1 && *first == '0') ||
              (d < eol && !mouse_is_ws(*d))) {
            result[t] = -((int64_t)(s - text) + 2);      // "+1", "01", "-0", "1.0", "C3": not canonical
            return;
          }
          ((int64_t *)outs[w])[row] = neg ? -v : v;
          q = d;
        }
        ++col;
      }
      if (col != ncol) {                                  // blank or ragged line
        result[t] = -((int64_t)(s - text) + 2);
        return;
      }
      s = eol < end ? eol + 1 : end;
    }
  };
  if (nthreads == 1) {
    parse_chunk(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(parse_chunk, t);
    for (auto &th : pool) th.join();
  }
  for (int t = 0; t < nthreads; ++t)
    if (result[t] < 0) return result[t];
  return nrows;
}

// Byte offset of every `ITEM: TIMESTEP` line of a dump (the frame index of ingest.index_lmp_dump): memmem over
// nthreads chunks of the (memory-mapped) file, matches kept when they open a line.  Returns the number of frames found;
// only the first `cap` offsets are written.
extern "C" int64_t mouse_index_dump(const char *text, int64_t len, int64_t *out, int64_t cap, int32_t nthreads) {
  static const char pat[] = "ITEM: TIMESTEP";
  const int64_t plen = (int64_t)sizeof(pat) - 1;
  if (!text || len < plen) return 0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if (len < (int64_t)1 << 22) nthreads = 1;
  std::vector<std::vector<int64_t>> found(nthreads);
  auto scan = [&](int t) {
    // matches that START in [a, b): the window runs plen - 1 bytes past b
    const int64_t a = len * t / nthreads, b = len * (t + 1) / nthreads;
    const int64_t wend = b + plen - 1 < len ? b + plen - 1 : len;
    const char *p = text + a, *end = text + wend;
    while (p < end) {
      const char *q = (const char *)memmem(p, (size_t)(end - p), pat, (size_t)plen);
      if (!q) break;
      if (q == text || q[-1] == '\n') found[t].push_back((int64_t)(q - text));
      p = q + plen;
    }
  };
  if (nthreads == 1) {
    scan(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(scan, t);
    for (auto &th : pool) th.join();
  }
  int64_t n = 0;
  for (int t = 0; t < nthreads; ++t)
    for (int64_t v : found[t]) {
      if (n < cap && out) out[n] = v;
      ++n;
    }
  return n;
}

// The ATOMS block of a LAMMPS dump frame (lammpack_types.py:377-395) straight into the arrays the pipeline needs:
// `nrows` lines of `ncol` whitespace-separated numbers; column col_id -> ids (int64), columns col_xyz[0..2] ->
// xyz[row][3] with the coordinate transform of the reference applied (mode 0 "scaled": lo + (hi - lo) * x, mode 1
// "cut": lo + x, mode 2 "uncut": x - un-fused, like the NumPy expressions), optional columns col_img[0..2] ->
// img[row][3] (int32).  Lines are distributed over `nthreads` host threads (cut at newlines).  Returns nrows, or
// -(byte offset + 1) of the first line that is short, long or holds a token that is not a number; -1 - len when the
// block does not hold exactly nrows lines.
extern "C" int64_t mouse_parse_dump_atoms(const char *text, int64_t len, int64_t nrows, int32_t ncol, int32_t col_id,
                                          const int32_t *col_xyz, const int32_t *col_img, int32_t mode,
                                          const double *lo, const double *hi, int64_t *ids, double *xyz, int32_t *img,
                                          int32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  if (len < (int64_t)1 << 16) nthreads = 1;
  std::vector<int64_t> cut(nthreads + 1), lines(nthreads, 0), result(nthreads, 0);
  cut[0] = 0;
  cut[nthreads] = len;
  for (int t = 1; t < nthreads; ++t) {
    int64_t c = len * t / nthreads;
    if (c < cut[t - 1]) c = cut[t - 1];
    while (c < len && text[c] != '\n') ++c;
    if (c < len) ++c;                                  // a chunk starts right behind a newline
    cut[t] = c;
  }
  auto blank = [&](int64_t a, int64_t b) {
    for (int64_t k = a; k < b; ++k)
      if (!mouse_is_ws(text[k])) return false;
    return true;
  };
  auto count_lines = [&](int t) {
    int64_t n = 0, a = cut[t];
    for (int64_t k = cut[t]; k < cut[t + 1]; ++k)
      if (text[k] == '\n') {
        n += !blank(a, k);
        a = k + 1;
      }
    if (a < cut[t + 1]) n += !blank(a, cut[t + 1]);
    lines[t] = n;
  };
  const double span[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
  auto parse_chunk = [&](int t, int64_t row) {
    const char *s = text + cut[t], *end = text + cut[t + 1];
    while (s < end) {
      const char *eol = (const char *)memchr(s, '\n', (size_t)(end - s));
      if (!eol) eol = end;
      const char *q = s;
      int col = 0;
      bool any = false;
      for (;;) {
        while (q < eol && mouse_is_ws(*q)) ++q;
        if (q >= eol) break;
        any = true;
        double v;
        const char *next = mouse_parse_token(q, eol, &v);
        if (!next || col >= ncol) {
          result[t] = -((int64_t)(q - text) + 1);
          return;
        }
        if (col == col_id) ids[row] = (int64_t)v;
        for (int a = 0; a < 3; ++a) {
          if (col == col_xyz[a]) {
            volatile double p = mode == 0 ? span[a] * v : v;      // (hi - lo) * x, rounded, then lo + ...
            xyz[3 * row + a] = mode == 2 ? v : lo[a] + p;
          }
          if (img && col == col_img[a]) img[3 * row + a] = (int32_t)v;
        }
        ++col;
        q = next;
      }
      if (any) {
        if (col != ncol) {
          result[t] = -((int64_t)(s - text) + 1);
          return;
        }
        ++row;
      }
      s = eol < end ? eol + 1 : end;
    }
  };
  if (nthreads == 1) {
    count_lines(0);
    if (lines[0] != nrows) return -1 - len;
    parse_chunk(0, 0);
    return result[0] < 0 ? result[0] : nrows;
  }
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(count_lines, t);
    for (auto &th : pool) th.join();
  }
  std::vector<int64_t> first(nthreads + 1, 0);
  for (int t = 0; t < nthreads; ++t) first[t + 1] = first[t] + lines[t];
  if (first[nthreads] != nrows) return -1 - len;
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(parse_chunk, t, first[t]);
    for (auto &th : pool) th.join();
  }
  for (int t = 0; t < nthreads; ++t)
    if (result[t] < 0) return result[t];
  return nrows;
}
