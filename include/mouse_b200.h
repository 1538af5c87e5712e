/* mouse_b200 - C-ABI of the B200-native local-ordering hot path.
 *
 * This is the drop-in boundary.  The reference (mglagolev/mouse) has no
 * FFI of its own: its seam is the set of plain Python functions in ordering_functions.py that
 * calculate_local_ordering.py / calculate_rdfs.py star-import.  Each entry point below names
 * the reference lines it replaces; the Python mirror (mouse_b200/ordering_functions.py) binds
 * them with ctypes and INTEGRATION.md shows the stub a maintainer would add upstream.
 *
 * Conventions: plain pointers and sizes only (no torch / C++ types); pointers marked DEV are
 * device pointers (e.g. tensor.data_ptr()), HOST are host pointers; `stream` is a cudaStream_t
 * passed as void*; every function returns 0 on success or a negative MOUSE_E* code and never
 * throws; nothing allocates behind the caller's back: scratch memory is a caller-provided
 * workspace sized by the *_workspace_bytes() queries.  All launches are asynchronous on
 * `stream`; entry points keep no mutable static state besides the thread-local last-error string (the
 * only cache is a table of immutable per-device facts: SM count, resident CTAs per SM).
 */
#ifndef MOUSE_B200_H
#define MOUSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOUSE_OK 0
#define MOUSE_EINVAL (-1)     /* bad argument */
#define MOUSE_ECUDA (-2)      /* CUDA runtime error (see mouse_last_error) */
#define MOUSE_EWORKSPACE (-3) /* workspace too small */

/* flags for the P2 / RDF sweeps */
#define MOUSE_SAME_MOLECULE 1u /* include pairs within one molecule (reference: sameMolecule=True) */
#define MOUSE_FORCE_EXACT 2u   /* use the all-FP64 sweep even when the FP32-filter sweep applies */
#define MOUSE_NO_CUTOFF 4u     /* reference: rcut < 0 (ordering_functions.py:102-103) */
#define MOUSE_FORCE_FAST 16u   /* P2: keep the FP32-filter sweep even when the cells are crowded (tests) */
/* Tuning fields inside `flags` (0 = automatic; there is no process-global tuning state): register-resident candidate
 * slots per lane of the fast P2 sweep (4, 6 or 8) and resident sweep warps per SM. */
#define MOUSE_SLOTS(n) (((uint32_t)(n) & 0xffu) << 8)
#define MOUSE_WARPS(n) (((uint32_t)(n) & 0xffu) << 16)

/* Cell grid chosen on the host for one frame (box, cutoff). */
typedef struct mouse_grid {
  double box[3];      /* orthorhombic edge lengths (reference Config.box(), lammpack_types.py:154) */
  double rcut;        /* cutoff; < 0 = none */
  double rc2;         /* the caller's Python `rcut**2` - threshold bits equal the reference's (:102) */
  double cell_w[3];   /* cell edge per axis, >= rcut*(1+1e-6) whenever ncell_axis >= 2 */
  int32_t ncell_axis[3];
  int32_t ncell;      /* product */
  int32_t fast;       /* 1: >= 3 cells on every axis and guard band small -> FP32-filter sweep valid */
  float guard;        /* relative half-width of the FP32 guard band around rc2 */
  /* Triclinic boxes (LAMMPS convention a = (lx,0,0), b = (xy,ly,0), c = (xz,yz,lz)): BUILD-DEFINED semantics, the
   * reference is orthorhombic only (lammpack_types.py:440-451 drops the tilt factors).  box[] holds lx, ly, lz; the
   * cells are parallelepipeds of the lattice (cell_w = edge / cells per axis); the minimum image of a displacement is
   * its representative in the brick |dx| <= lx/2, |dy| <= ly/2, |dz| <= lz/2 found axis by axis from z to x with the
   * reference's own step dr - L*around(dr/L) (ordering_functions.py:98-100), removing the tilt components of the
   * subtracted lattice vector from the lower axes.  tilt = 0 reproduces the orthorhombic results bit for bit. */
  double tilt[3];     /* xy, xz, yz */
  int32_t triclinic;  /* 1: planned by mouse_grid_plan_triclinic (the tilt-aware code paths run even for tilt = 0) */
  int32_t reserved;
} mouse_grid_t;

/* Totals written by the fused reductions (device or host memory, see each function). */
typedef struct mouse_p2_totals {
  double sum_mean;       /* sum_i m_i over references with >= 1 counted neighbour (:205) */
  int64_t n_refs;        /* i_s (:206) */
  int64_t n_pairs;       /* sum_i count_i = ordered in-cutoff pairs the reference counts */
  int64_t n_zero_len;    /* zero-length bonds in the neighbour set (NaN poisoning, :112) */
  int64_t n_hist_over;   /* values with bin index >= nbins (reference raises IndexError) */
  int64_t n_band;        /* candidate pairs that needed the FP64 cutoff re-check (diagnostic) */
} mouse_p2_totals_t;

const char *mouse_version(void);
const char *mouse_last_error(void);

/* HOST only.  Chooses the cell grid for `n_items` points.  Returns MOUSE_OK. */
int mouse_grid_plan(const double *box /*HOST[3]*/, double rcut, double rc2, int64_t n_items,
                    mouse_grid_t *grid /*HOST out*/);

/* HOST only.  The same for a triclinic box (tilt = {xy, xz, yz}); cells per axis come from the perpendicular widths
 * of the box, `fast` additionally needs every cell plus the cutoff to fit half the box edge in Cartesian extent. */
int mouse_grid_plan_triclinic(const double *box /*HOST[3]*/, const double *tilt /*HOST[3]*/, double rcut, double rc2,
                              int64_t n_items, mouse_grid_t *grid /*HOST out*/);

/* K1.  Replaces vector_and_center_pbc_trim method 1 (lammpack_types.py:587-599) as called per
 * bond by createNumpyBondsArrayFromConfig (ordering_functions.py:29) and again per reference
 * bond (:92).  vec/ctr are the rows 3-5 / 6-8 of the reference's (9,N_b) npBonds array:
 * SoA, vec[a*n_bonds + b]. */
int mouse_bond_build(const double *pos /*DEV [n_atoms][3]*/, const int32_t *bond_atoms /*DEV [n_bonds][2]*/,
                     int64_t n_bonds, const double *box /*HOST[3]*/, double *vec /*DEV [3][n_bonds]*/,
                     double *ctr /*DEV [3][n_bonds]*/, void *stream);

/* K1 for a triclinic box: the np.remainder form of lammpack_types.py:587-599 axis by axis from z to x (build-defined). */
int mouse_bond_build_triclinic(const double *pos, const int32_t *bond_atoms, int64_t n_bonds, const double *box /*HOST[3]*/,
                               const double *tilt /*HOST[3]*/, double *vec, double *ctr, void *stream);

/* Workspace for cell list + sweep on `n_items` bonds (or atoms) with `grid`. */
size_t mouse_workspace_bytes(int64_t n_items, const mouse_grid_t *grid);

/* K2.  No reference counterpart (the reference is all-pairs, ordering_functions.py:94-115);
 * builds the cell list over the bonds selected by nbr_mask (createNumpyBondsArrayFromConfig's
 * allowedBondTypes filter, :25-28): single-digit radix (counting) sort by cell id, per-cell
 * offsets, canonical (ascending bond index) order inside each cell, payload gathered into
 * sorted float4 / double4 arrays inside the workspace.
 * mol: per-bond molecule code (hash8(atom1.mol_id), :31,39), ref_mask: bonds that act as
 * references (:198-199).  Masks may be NULL (= all ones). */
int mouse_cell_build(const double *vec, const double *ctr, const int32_t *mol /*DEV [n]*/,
                     const uint8_t *nbr_mask /*DEV [n] or NULL*/, const uint8_t *ref_mask /*DEV [n] or NULL*/,
                     int64_t n_bonds, const mouse_grid_t *grid /*HOST*/, void *ws /*DEV*/, size_t ws_bytes,
                     void *stream);

/* K3.  Replaces calculateAveCosSqForReference (ordering_functions.py:91-119) for every
 * reference bond at once: out_sum[b] = sum of cos^2 over the counted neighbours of bond b,
 * out_cnt[b] = their number (0 = the reference raised NameError / bond is not a reference).
 * Outputs are indexed by ORIGINAL bond index and fully overwritten. */
int mouse_p2_sweep(const double *vec, const double *ctr, int64_t n_bonds, const mouse_grid_t *grid,
                   uint32_t flags, void *ws /*DEV, holds the cell list*/, size_t ws_bytes,
                   double *out_sum /*DEV [n]*/, int32_t *out_cnt /*DEV [n]*/, void *stream);

/* Optional second pass: the counted neighbour lists themselves (CSR over original bond index,
 * rows ascending).  row_ptr = exclusive scan of out_cnt (n+1 entries, int64), computed here; entries at or beyond
 * `capacity` (the allocated length of nbr_idx, normally sum(cnt)) are never written. */
int mouse_p2_pair_lists(const double *vec, const double *ctr, int64_t n_bonds, const mouse_grid_t *grid,
                        uint32_t flags, void *ws, size_t ws_bytes, const int32_t *cnt /*DEV [n]*/,
                        int64_t *row_ptr /*DEV [n+1] out*/, int32_t *nbr_idx /*DEV [capacity] out*/,
                        int64_t capacity, void *stream);

/* K4.  Replaces the tail of CalculateOrientationOrderParameter (ordering_functions.py:202-206,
 * 221-224) and updateHistogram (lammpack_misc.py:185-189): per-vector mean cos^2 (NaN when
 * count == 0), totals, optional histogram of 1.5*m-0.5 with the reference's bin rule.
 * hist may be NULL.  totals is DEV memory. */
int mouse_p2_reduce(const double *sum /*DEV [n]*/, const int32_t *cnt /*DEV [n]*/, int64_t n_bonds,
                    const mouse_grid_t *grid, void *ws, size_t ws_bytes, double *out_mean /*DEV [n]*/,
                    mouse_p2_totals_t *totals /*DEV*/, int64_t *hist /*DEV [nbins] or NULL*/, double smin, double smax,
                    int32_t nbins, void *stream);

/* K1..K4 in one call, device buffers (the timed hot path of bench.py `value`). */
int mouse_p2_frame(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                   const uint8_t *ref_mask, int64_t n_bonds, const mouse_grid_t *grid, uint32_t flags,
                   void *ws, size_t ws_bytes, double *vec, double *ctr, double *sum, int32_t *cnt,
                   double *mean, mouse_p2_totals_t *totals, void *stream);

/* The same frame captured ONCE into a CUDA graph (trajectories: fixed topology, box, cutoff and buffers - only the
 * contents of `pos` change): _create runs no kernel, it records mouse_p2_frame's launch train; _launch replays it on
 * `stream` with one driver call per frame; _nodes = kernels + memsets recorded; _destroy frees the handle.  The
 * buffers named at creation must stay alive and must not be shared with a graph launched concurrently. */
int mouse_p2_frame_graph_create(const double *pos, const int32_t *bond_atoms, const int32_t *mol,
                                const uint8_t *nbr_mask, const uint8_t *ref_mask, int64_t n_bonds,
                                const mouse_grid_t *grid, uint32_t flags, void *ws, size_t ws_bytes, double *vec,
                                double *ctr, double *sum, int32_t *cnt, double *mean, mouse_p2_totals_t *totals,
                                void **out_graph /*HOST out*/);
int mouse_p2_frame_graph_launch(void *graph, void *stream);
int mouse_p2_frame_graph_nodes(void *graph);
int mouse_p2_frame_graph_destroy(void *graph);

/* mouse_p2_frame with two optional cudaEvent_t handles (NULL = none) recorded on `stream` right before and right
 * after the K3 launches: lets a caller time the pair sweep inside a frame (bench.py's roofline line). */
int mouse_p2_frame_timed(const double *pos, const int32_t *bond_atoms, const int32_t *mol, const uint8_t *nbr_mask,
                         const uint8_t *ref_mask, int64_t n_bonds, const mouse_grid_t *grid, uint32_t flags,
                         void *ws, size_t ws_bytes, double *vec, double *ctr, double *sum, int32_t *cnt,
                         double *mean, mouse_p2_totals_t *totals, void *ev_sweep_begin, void *ev_sweep_end,
                         void *stream);

/* K5.  Replaces calculateRdfForReference (ordering_functions.py:66-88) summed over all atoms as
 * in CalculatePairCorrelationFunctions (:162-173), accumulating onto zeros.
 * type: per-atom type code in [0,ntypes) or < 0 (not selected); edges: NumPy's bin edges
 * (nbin+1 doubles, HOST); hist: int64 [ntypes*(ntypes+1)/2][nbin], overwritten. */
int mouse_rdf_frame(const double *pos /*DEV [n][3]*/, const int32_t *type /*DEV [n]*/, const int32_t *mol /*DEV [n]*/,
                    int64_t n_atoms, int32_t ntypes, const mouse_grid_t *grid, uint32_t flags,
                    const double *edges /*HOST [nbin+1]*/, int32_t nbin, void *ws, size_t ws_bytes,
                    int64_t *hist /*DEV*/, void *stream);

/* mouse_rdf_frame with two optional cudaEvent_t handles recorded right before / after the K5 launches (bench.py). */
int mouse_rdf_frame_timed(const double *pos, const int32_t *type, const int32_t *mol, int64_t n_atoms, int32_t ntypes,
                          const mouse_grid_t *grid, uint32_t flags, const double *edges, int32_t nbin, void *ws,
                          size_t ws_bytes, int64_t *hist, void *ev_sweep_begin, void *ev_sweep_end, void *stream);

/* One reference against a caller-built array: the reference's per-call granularity.
 * mouse_p2_single_ref replaces calculateAveCosSqForReference (ordering_functions.py:91-119):
 *   np_bonds DEV (9,n) f64 row-major = createNumpyBondsArrayFromConfig's array (:41);
 *   ref HOST[8] = {bRef.x,y,z, bRefCenter.x,y,z, refBond.num, hash8(refBond.atom1.mol_id)};
 *   rcut < 0 = no cutoff; flags: MOUSE_SAME_MOLECULE, MOUSE_INCLUDE_SELF (excludeSelf=False);
 *   out DEV[2] = {sum of unmasked cos^2, unmasked count} (count == 0 -> the reference raises
 *   NameError("No values to calculate"), :117).  A zero-length bRef is the caller's check (:93,119).
 * mouse_rdf_single_ref replaces calculateRdfForReference (:66-88):
 *   np_atoms DEV (6,n) f64 = createNumpyAtomsArrayFromConfig's array (:63);
 *   ref HOST[5] = {x, y, z, refAtom.num, hash8(refAtom.mol_id)}; edges HOST[nbin+1] (NumPy's);
 *   hist DEV[nbin+1]: bins, then the number of unmasked distances (0 -> NameError, :88). */
#define MOUSE_INCLUDE_SELF 8u
size_t mouse_single_ref_workspace_bytes(void);
int mouse_p2_single_ref(const double *np_bonds, int64_t n, const double *ref, const double *box, double rcut,
                        double rc2, uint32_t flags, void *ws, double *out, void *stream);
int mouse_rdf_single_ref(const double *np_atoms, int64_t n, const double *ref, const double *box, uint32_t flags,
                         const double *edges, int32_t nbin, void *ws, int64_t *hist, void *stream);

/* P2 of every bond against one director: replaces the per-bond loop of CalculateCrystallinityParameter's histo branch
 * (ordering_functions.py:263-273).  vec = K1's SoA [3][n]; mask NULL = all bonds; director HOST[3] (ave_b, :242-251);
 * out_s DEV[n] = 1.5*cos^2-0.5 per bond (NaN where masked out); hist DEV[nbins+1]: bins filled with the rule of
 * updateHistogram (lammpack_misc.py:185-189), hist[nbins] = values the reference would raise IndexError on. */
int mouse_director_p2(const double *vec, const uint8_t *mask, int64_t n_bonds, const double *director, double *out_s,
                      int64_t *hist, double smin, double smax, int32_t nbins, void *stream);

/* Per-atom colouring of --out-pdb (storeAsAtomtypes, ordering_functions.py:207-220): for every atom the mean of
 * 1.5*m-0.5 over its incident bonds with cnt > 0 (CSR adjacency row_ptr[n_atoms+1] / bond_of[], bonds ascending),
 * as the integer of the type string: code >= 0 -> "C%02d" % code; code < 0 -> "Cm%02d" % (-code - 1); INT32_MIN -> the
 * atom has no bond with a value and keeps its type.  direct != 0: `mean` already holds the per-bond value (director
 * P2 of CalculateCrystallinityParameter, :263-281) and cnt is a 0/1 selection.  All pointers DEV. */
int mouse_atom_colour(const double *mean, const int32_t *cnt, const int64_t *row_ptr, const int32_t *bond_of,
                      int64_t n_atoms, int32_t direct, int32_t *code, void *stream);

/* Debug: the 8 int64 workspace counters (0 zero-length, 1 guard-band redos, 4-7 phase cycle timers when built
 * with -DMOUSE_PHASE_TIMERS); synchronises `stream`. */
int mouse_debug_counters(const mouse_grid_t *grid, int64_t n_items, const void *ws, int64_t *out8 /*HOST*/, void *stream);

/* HOST only (file ingestion row).  Parses the whitespace-separated decimal numbers of a text
 * block (the ATOMS section of a LAMMPS dump, lammpack_types.py:377-395 reads it with one float() per token) into
 * out[0..cap).  Correctly rounded: numbers with <= 15 significant digits and a decimal exponent within +-22 take the
 * exact fast path (integer mantissa times / divided by an exactly representable power of ten, one rounding), all
 * others go through strtod.  Returns the number of values written, or -(offset + 1) of the first token that is not a
 * number or does not fit. */
int64_t mouse_parse_doubles(const char *text /*HOST*/, int64_t len, double *out /*HOST*/, int64_t cap);
/* The same on `nthreads` host threads (the text is cut at whitespace, tokens are counted, then parsed in place):
 * identical values and error position. */
int64_t mouse_parse_doubles_mt(const char *text /*HOST*/, int64_t len, double *out /*HOST*/, int64_t cap,
                               int32_t nthreads);

/* The ATOMS block of one LAMMPS dump frame straight into arrays (lammpack_types.py:377-395: one float() per token, then
 * `lo + (hi - lo) * xs` for scaled coordinates): nrows lines of ncol numbers; col_id -> ids, col_xyz[3] -> xyz[row][3]
 * with the reference's coordinate transform (mode 0 scaled, 1 cut: lo + x, 2 uncut: x; un-fused), col_img[3] (or NULL)
 * -> img[row][3].  Lines are split over nthreads host threads.  Returns nrows or a negative error (see api.cu). */
int64_t mouse_parse_dump_atoms(const char *text /*HOST*/, int64_t len, int64_t nrows, int32_t ncol, int32_t col_id,
                               const int32_t *col_xyz, const int32_t *col_img, int32_t mode, const double *lo,
                               const double *hi, int64_t *ids, double *xyz, int32_t *img, int32_t nthreads);

/* HOST only.  One section of a LAMMPS data file (the per-line readers AtomFromLmpDataLine / BondFromLmpDataLine,
 * lammpack_lmp_functions.py:3-33, called from lammpack_types.py:400-464): nrows lines of exactly ncol tokens; column
 * cols[w] -> outs[w][row] as double (kinds[w] == 0, correctly rounded) or as int64 (kinds[w] == 1, canonical integers
 * only, so that the token can be printed back: types and molecule ids are strings in the reference).  *consumed = bytes
 * of the block.  Returns nrows, or < 0 when the block has another shape (the caller takes its line-by-line path). */
int64_t mouse_parse_table(const char *text /*HOST*/, int64_t len, int64_t nrows, int32_t ncol, int32_t nwant,
                          const int32_t *cols, const int32_t *kinds, void *const *outs, int64_t *consumed,
                          int32_t nthreads);

/* HOST only.  Frame index of a LAMMPS dump (lammpack_types.py:344-398 walks the file line by line looking for
 * `ITEM: TIMESTEP`): byte offset of every line that starts with it, ascending, found with nthreads host threads.
 * Returns the number of frames; only the first `cap` offsets are written to out (call again with a larger array). */
int64_t mouse_index_dump(const char *text /*HOST*/, int64_t len, int64_t *out /*HOST*/, int64_t cap, int32_t nthreads);

/* The same block parsed ON THE DEVICE: `text` (DEV, 16-byte aligned, len < 2^31) holds the raw bytes of the ATOMS block
 * as they lie in the file; ids / xyz / img (DEV) receive what mouse_parse_dump_atoms would write, bit for bit (the exact
 * fast path of the number parser: <= 15 significant digits, decimal exponent within +-22, one IEEE rounding; the
 * reference's coordinate transform un-fused).  row_of_line (DEV int32[nrows] or NULL): xyz / img of line k go to row
 * row_of_line[k] (the id -> topology row permutation of a previous frame); ids are always written in line order.
 * expect_ids (DEV or NULL): ids of the previous frame, mismatches are counted.  lo / hi / col_*: HOST.
 * status (DEV int64[4], written by the call): [0] non-blank lines found (must equal nrows), [1] byte offset of the
 * first line with a wrong number of columns (INT64_MAX: none), [2] tokens that need the host parser (long mantissas,
 * nan / inf, malformed), [3] lines whose id differs from expect_ids.  Unless [0] == nrows, [1] == INT64_MAX and
 * [2] == 0 the outputs are incomplete and the caller parses the frame with mouse_parse_dump_atoms, which is
 * authoritative for errors.  ws: mouse_parse_dump_ws_bytes(len, nrows) bytes (DEV).  Asynchronous on `stream`. */
size_t mouse_parse_dump_ws_bytes(int64_t len, int64_t nrows);
int mouse_parse_dump_atoms_dev(const char *text /*DEV*/, int64_t len, int64_t nrows, int32_t ncol, int32_t col_id,
                               const int32_t *col_xyz /*HOST*/, const int32_t *col_img /*HOST or NULL*/, int32_t mode,
                               const double *lo /*HOST*/, const double *hi /*HOST*/, const int32_t *row_of_line,
                               const int64_t *expect_ids, int64_t *ids, double *xyz, int32_t *img, void *ws,
                               size_t ws_bytes, int64_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MOUSE_B200_H */
