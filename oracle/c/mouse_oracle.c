/* TEST INFRASTRUCTURE ONLY - plain-C restatement of the reference's local-P2 and RDF
 * arithmetic on a CPU cell list.  Not linked into, nor called by, the product library.
 *
 * Parity pin: validated in tests/test_oracle_golden.py against oracle/p2_oracle.py (which is
 * itself pinned bit-for-bit to golden vectors produced by the real reference) and directly
 * against the golden vectors.  Build with -ffp-contract=off and WITHOUT -ffast-math: every
 * floating-point expression below must evaluate un-fused, in the order written, exactly like
 * the NumPy expressions it restates.
 *
 * Reference lines restated (paths relative to the reference root):
 *   bond vector / midpoint     lammpack_types.py:587-599  (np.remainder form)
 *   per-pair P2 terms          ordering_functions.py:95-115
 *   per-pair RDF distance      ordering_functions.py:69-86 + numpy.histogram uniform-bin rule
 *
 * The reference visits ALL pairs; a cell list only prunes candidates that the reference's own
 * `rsq <= rcut**2` test would reject, so counted sets are identical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* np.remainder(a, b) for b > 0: fmod, then shift negative results by +b (npy_divmod) */
static double np_remainder(double a, double b) {
  double m = fmod(a, b);
  if (m != 0.0) {
    if (m < 0.0) m += b;
  } else {
    m = copysign(0.0, b);
  }
  return m;
}

/* ---- triclinic boxes: BUILD-DEFINED semantics, parity unpinned (the reference is orthorhombic only,
 * lammpack_types.py:440-451 reads no tilt factors).  LAMMPS convention: a = (lx,0,0), b = (xy,ly,0), c = (xz,yz,lz),
 * tilt = {xy, xz, yz}.  The minimum image of a displacement is its representative in the brick |dx| <= lx/2,
 * |dy| <= ly/2, |dz| <= lz/2, found axis by axis from z to x with the reference's own `dr - L*around(dr/L)` step;
 * the tilt components of the subtracted lattice vector are removed from the lower axes.  With tilt = 0 every
 * expression below reduces bit for bit to the orthorhombic one. */
static inline void tri_min_image(double *dx, double *dy, double *dz, const double *box, const double *tilt) {
  double nz = rint(*dz / box[2]);
  *dz = *dz - box[2] * nz;
  *dy = *dy - tilt[2] * nz;
  *dx = *dx - tilt[1] * nz;
  double ny = rint(*dy / box[1]);
  *dy = *dy - box[1] * ny;
  *dx = *dx - tilt[0] * ny;
  double nx = rint(*dx / box[0]);
  *dx = *dx - box[0] * nx;
}

/* bond vector / midpoint, triclinic: the np.remainder form of lammpack_types.py:587-599 axis by axis from z to x */
void mo_bond_build_tri(const double *pos, const int32_t *bonds, int64_t n_bonds, const double *box,
                       const double *tilt, double *vec, double *ctr) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < n_bonds; ++b) {
    const double *p1 = pos + 3 * (int64_t)bonds[2 * b];
    const double *p2 = pos + 3 * (int64_t)bonds[2 * b + 1];
    double x[3], xc[3], w[3], t[3];
    for (int a = 0; a < 3; ++a) {
      x[a] = p2[a] - p1[a];
      xc[a] = (p2[a] + p1[a]) / 2.;
      w[a] = x[a];
    }
    t[2] = np_remainder(w[2] + box[2] / 2., box[2]) - box[2] / 2.;
    double nz = rint((w[2] - t[2]) / box[2]);
    w[1] = w[1] - tilt[2] * nz;
    w[0] = w[0] - tilt[1] * nz;
    t[1] = np_remainder(w[1] + box[1] / 2., box[1]) - box[1] / 2.;
    double ny = rint((w[1] - t[1]) / box[1]);
    w[0] = w[0] - tilt[0] * ny;
    t[0] = np_remainder(w[0] + box[0] / 2., box[0]) - box[0] / 2.;
    for (int a = 0; a < 3; ++a) {
      vec[3 * b + a] = t[a];
      ctr[3 * b + a] = xc[a] + (t[a] - x[a]) / 2.;
    }
  }
}

/* lammpack_types.py:587-599; pos (n_atoms,3), bonds (n_bonds,2) 0-based, out vec/ctr (n_bonds,3) */
void mo_bond_build(const double *pos, const int32_t *bonds, int64_t n_bonds, const double *box,
                   double *vec, double *ctr) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < n_bonds; ++b) {
    const double *p1 = pos + 3 * (int64_t)bonds[2 * b];
    const double *p2 = pos + 3 * (int64_t)bonds[2 * b + 1];
    for (int a = 0; a < 3; ++a) {
      double L = box[a];
      double x = p2[a] - p1[a];
      double xc = (p2[a] + p1[a]) / 2.;
      double xt = np_remainder(x + L / 2., L) - L / 2.;
      vec[3 * b + a] = xt;
      ctr[3 * b + a] = xc + (xt - x) / 2.;
    }
  }
}

typedef struct {
  int n[3];
  double w[3];
  int64_t ncell;
  int64_t *start; /* ncell+1 */
  int64_t *items; /* sorted item ids */
} cells_t;

static int cell_axis(double c, double L, double w, int n) {
  double x = c - L * floor(c / L);
  if (x >= L) x -= L;
  if (x < 0.0) x = 0.0;
  int k = (int)(x / w);
  return k >= n ? n - 1 : k;
}

/* perpendicular widths of a triclinic box along a, b, c (distance between opposite faces) */
static void tri_widths(const double *box, const double *tilt, double *w) {
  double lx = box[0], ly = box[1], lz = box[2], xy = tilt[0], xz = tilt[1], yz = tilt[2];
  double bcx = ly * lz, bcy = -xy * lz, bcz = xy * yz - ly * xz; /* b x c */
  w[0] = lx * ly * lz / sqrt(bcx * bcx + bcy * bcy + bcz * bcz);
  w[1] = ly * lz / sqrt(lz * lz + yz * yz);                      /* V / |a x c| */
  w[2] = lz;
}

static int cells_build_t(cells_t *g, const double *xyz, int64_t n, const double *box, const double *tilt, double rc) {
  double wid[3] = {box[0], box[1], box[2]};
  if (tilt) tri_widths(box, tilt, wid);
  for (int a = 0; a < 3; ++a) {
    int k = 1;
    if (rc > 0.0) {
      double q = wid[a] / (rc * (1.0 + 1e-6));
      k = q >= 1.0 ? (q > 512.0 ? 512 : (int)q) : 1;
    }
    g->n[a] = k;
    g->w[a] = box[a] / k;
  }
  g->ncell = (int64_t)g->n[0] * g->n[1] * g->n[2];
  g->start = (int64_t *)calloc((size_t)g->ncell + 1, sizeof(int64_t));
  g->items = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
  int64_t *cid = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
  if (!g->start || !g->items || !cid) return 1;
  for (int64_t i = 0; i < n; ++i) {
    double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    if (tilt) { /* fractional coordinates times the edge length: cells are parallelepipeds of the lattice */
      double sz = z / box[2];
      double sy = (y - tilt[2] * sz) / box[1];
      double sx = (x - tilt[0] * sy - tilt[1] * sz) / box[0];
      x = sx * box[0], y = sy * box[1], z = sz * box[2];
    }
    int cx = cell_axis(x, box[0], g->w[0], g->n[0]);
    int cy = cell_axis(y, box[1], g->w[1], g->n[1]);
    int cz = cell_axis(z, box[2], g->w[2], g->n[2]);
    cid[i] = ((int64_t)cz * g->n[1] + cy) * g->n[0] + cx;
    g->start[cid[i] + 1]++;
  }
  for (int64_t c = 0; c < g->ncell; ++c) g->start[c + 1] += g->start[c];
  int64_t *cur = (int64_t *)malloc((size_t)g->ncell * sizeof(int64_t));
  memcpy(cur, g->start, (size_t)g->ncell * sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) g->items[cur[cid[i]]++] = i; /* stable: ascending id per cell */
  free(cur);
  free(cid);
  return 0;
}


static void cells_free(cells_t *g) {
  free(g->start);
  free(g->items);
}

/* neighbour cell indices along one axis: {c-1,c,c+1} wrapped when n >= 3, else every cell */
static int axis_neighbours(int c, int n, int out[3]) {
  if (n >= 3) {
    out[0] = (c + n - 1) % n;
    out[1] = c;
    out[2] = (c + 1) % n;
    return 3;
  }
  for (int k = 0; k < n; ++k) out[k] = k;
  return n;
}

static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

/* ordering_functions.py:91-119 for every reference bond.
 * vec, ctr: (n,3) float64; mol: (n,) int64 codes compared for equality only;
 * ref_mask / nbr_mask: (n,) uint8; rcut < 0 means "no cutoff" (rc2 ignored);
 * rc2 is the caller's Python `rcut**2` so that the threshold bits match the reference.
 * out_sum (n) double, out_cnt (n) int64.  If nbr_ptr != NULL it must hold n+1 entries and the
 * function returns the counted neighbour lists (ascending index) in a malloc'ed array through
 * *nbr_idx (free with mo_free).  Zero-length reference bonds are skipped (count 0); zero-length
 * neighbour bonds poison every mean with NaN and add to every count (0/0 in the reference).
 */
static int mo_p2_impl(const double *vec, const double *ctr, const int64_t *mol, const uint8_t *ref_mask,
                      const uint8_t *nbr_mask, int64_t n, const double *box, const double *tilt, double rcut,
                      double rc2, int same_molecule, double *out_sum, int64_t *out_cnt, int64_t *nbr_ptr,
                      int64_t **nbr_idx) {
  cells_t g;
  if (cells_build_t(&g, ctr, n, box, tilt, rcut)) return 1;
  int64_t n_zero_nbr = 0;
  uint8_t *zero = (uint8_t *)malloc((size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) {
    zero[i] = (vec[3 * i] == 0.0 && vec[3 * i + 1] == 0.0 && vec[3 * i + 2] == 0.0);
    if (zero[i] && nbr_mask[i]) n_zero_nbr++;
  }
  const double Lx = box[0], Ly = box[1], Lz = box[2];
  int64_t *lists = NULL;
  for (int pass = 0; pass < (nbr_ptr ? 2 : 1); ++pass) {
    if (pass == 1) {
      nbr_ptr[0] = 0;
      for (int64_t i = 0; i < n; ++i) nbr_ptr[i + 1] = nbr_ptr[i] + (zero[i] || !ref_mask[i] || !nbr_mask[i] ? 0 : out_cnt[i]);
      lists = (int64_t *)malloc((size_t)(nbr_ptr[n] > 0 ? nbr_ptr[n] : 1) * sizeof(int64_t));
      if (!lists) return 1;
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t c = 0; c < g.ncell; ++c) {
      if (g.start[c] == g.start[c + 1]) continue;
      int cx = (int)(c % g.n[0]), cy = (int)((c / g.n[0]) % g.n[1]), cz = (int)(c / ((int64_t)g.n[0] * g.n[1]));
      int nx[3], ny[3], nz[3];
      int mx = axis_neighbours(cx, g.n[0], nx), my = axis_neighbours(cy, g.n[1], ny), mz = axis_neighbours(cz, g.n[2], nz);
      for (int64_t ii = g.start[c]; ii < g.start[c + 1]; ++ii) {
        int64_t i = g.items[ii];
        if (!ref_mask[i] || !nbr_mask[i] || zero[i]) {
          if (pass == 0) { out_sum[i] = 0.0; out_cnt[i] = 0; }
          continue;
        }
        const double bix = vec[3 * i], biy = vec[3 * i + 1], biz = vec[3 * i + 2];
        const double cix = ctr[3 * i], ciy = ctr[3 * i + 1], ciz = ctr[3 * i + 2];
        const double nr = bix * bix + biy * biy + biz * biz;
        double s = 0.0;
        int64_t cnt = 0;
        int64_t *lp = pass == 1 ? lists + nbr_ptr[i] : NULL;
        for (int a = 0; a < mz; ++a)
          for (int b = 0; b < my; ++b)
            for (int d = 0; d < mx; ++d) {
              int64_t cc = ((int64_t)nz[a] * g.n[1] + ny[b]) * g.n[0] + nx[d];
              for (int64_t jj = g.start[cc]; jj < g.start[cc + 1]; ++jj) {
                int64_t j = g.items[jj];
                if (j == i || !nbr_mask[j] || zero[j]) continue;
                if (!same_molecule && mol[j] == mol[i]) continue;
                if (rcut >= 0.0) {
                  double drx = ctr[3 * j] - cix, dry = ctr[3 * j + 1] - ciy, drz = ctr[3 * j + 2] - ciz;
                  double tx, ty, tz;
                  if (tilt) {
                    tx = drx, ty = dry, tz = drz;
                    tri_min_image(&tx, &ty, &tz, box, tilt);
                  } else {
                    tx = drx - Lx * rint(drx / Lx);
                    ty = dry - Ly * rint(dry / Ly);
                    tz = drz - Lz * rint(drz / Lz);
                  }
                  double rsq = tx * tx + ty * ty + tz * tz;
                  if (!(rsq <= rc2)) continue;
                }
                double bjx = vec[3 * j], bjy = vec[3 * j + 1], bjz = vec[3 * j + 2];
                double dot = bix * bjx + biy * bjy + biz * bjz;
                double nj = bjx * bjx + bjy * bjy + bjz * bjz;
                double cs = dot * dot / nr / nj;
                if (cs != 0.0) {
                  if (pass == 0) { s += cs; cnt++; }
                  else *lp++ = j;
                }
              }
            }
        if (pass == 0) {
          if (n_zero_nbr) { s = NAN; }
          out_sum[i] = s;
          out_cnt[i] = cnt;
        } else {
          qsort(lists + nbr_ptr[i], (size_t)out_cnt[i], sizeof(int64_t), cmp_i64);
        }
      }
    }
  }
  if (n_zero_nbr)
    for (int64_t i = 0; i < n; ++i)
      if (ref_mask[i] && nbr_mask[i] && !zero[i]) out_cnt[i] += n_zero_nbr;
  if (nbr_ptr) *nbr_idx = lists;
  free(zero);
  cells_free(&g);
  return 0;
}

int mo_p2(const double *vec, const double *ctr, const int64_t *mol, const uint8_t *ref_mask,
          const uint8_t *nbr_mask, int64_t n, const double *box, double rcut, double rc2,
          int same_molecule, double *out_sum, int64_t *out_cnt, int64_t *nbr_ptr, int64_t **nbr_idx) {
  return mo_p2_impl(vec, ctr, mol, ref_mask, nbr_mask, n, box, NULL, rcut, rc2, same_molecule, out_sum, out_cnt,
                    nbr_ptr, nbr_idx);
}

/* triclinic variant (build-defined minimum image, see tri_min_image); tilt = {xy, xz, yz} */
int mo_p2_tri(const double *vec, const double *ctr, const int64_t *mol, const uint8_t *ref_mask,
              const uint8_t *nbr_mask, int64_t n, const double *box, const double *tilt, double rcut, double rc2,
              int same_molecule, double *out_sum, int64_t *out_cnt, int64_t *nbr_ptr, int64_t **nbr_idx) {
  return mo_p2_impl(vec, ctr, mol, ref_mask, nbr_mask, n, box, tilt, rcut, rc2, same_molecule, out_sum, out_cnt,
                    nbr_ptr, nbr_idx);
}

void mo_free(void *p) { free(p); }

/* numpy.histogram (uniform bins) index rule, NumPy 2.3 _histograms_impl.histogram */
static inline int64_t np_bin(double d, double first, double last, int64_t nbin, const double *edges) {
  if (!(d >= first) || !(d <= last)) return -1;
  double f = ((d - first) / (last - first)) * (double)nbin;
  int64_t k = (int64_t)f;
  if (k == nbin) k -= 1;
  if (d < edges[k]) k -= 1;
  if (d >= edges[k + 1] && k != nbin - 1) k += 1;
  return k;
}

/* ordering_functions.py:66-88 summed as in :162-173, counts accumulated onto zeros.
 * pos (n,3); type (n,) int32 in [0,ntypes) or <0 for "not selected"; mol (n,) int64;
 * edges (nbin+1) = numpy's bin edges for range (rmin,rmax); hist [(ntypes*(ntypes+1)/2), nbin]
 * int64, pair slot of (a<=b) = a*ntypes - a*(a-1)/2 + (b-a).  rcell = cell edge to use (rmax).
 * Every ORDERED pair (i reference, j neighbour) adds one count (so unordered pairs count twice).
 */
static int mo_rdf_impl(const double *pos, const int32_t *type, const int64_t *mol, int64_t n, int ntypes,
                       const double *box, const double *tilt, double rcell, const double *edges, int64_t nbin,
                       int same_molecule, int64_t *hist) {
  cells_t g;
  if (cells_build_t(&g, pos, n, box, tilt, rcell)) return 1;
  const double Lx = box[0], Ly = box[1], Lz = box[2];
  const double first = edges[0], last = edges[nbin];
  int64_t npair = (int64_t)ntypes * (ntypes + 1) / 2;
  memset(hist, 0, (size_t)(npair * nbin) * sizeof(int64_t));
#pragma omp parallel
  {
    int64_t *h = (int64_t *)calloc((size_t)(npair * nbin), sizeof(int64_t));
#pragma omp for schedule(dynamic, 16)
    for (int64_t c = 0; c < g.ncell; ++c) {
      int cx = (int)(c % g.n[0]), cy = (int)((c / g.n[0]) % g.n[1]), cz = (int)(c / ((int64_t)g.n[0] * g.n[1]));
      int nx[3], ny[3], nz[3];
      int mx = axis_neighbours(cx, g.n[0], nx), my = axis_neighbours(cy, g.n[1], ny), mz = axis_neighbours(cz, g.n[2], nz);
      for (int64_t ii = g.start[c]; ii < g.start[c + 1]; ++ii) {
        int64_t i = g.items[ii];
        if (type[i] < 0) continue;
        for (int a = 0; a < mz; ++a)
          for (int b = 0; b < my; ++b)
            for (int d = 0; d < mx; ++d) {
              int64_t cc = ((int64_t)nz[a] * g.n[1] + ny[b]) * g.n[0] + nx[d];
              for (int64_t jj = g.start[cc]; jj < g.start[cc + 1]; ++jj) {
                int64_t j = g.items[jj];
                if (j == i || type[j] < 0) continue;
                if (!same_molecule && mol[j] == mol[i]) continue;
                double drx = pos[3 * j] + -1. * pos[3 * i], dry = pos[3 * j + 1] + -1. * pos[3 * i + 1],
                       drz = pos[3 * j + 2] + -1. * pos[3 * i + 2];
                double tx, ty, tz;
                if (tilt) {
                  tx = drx, ty = dry, tz = drz;
                  tri_min_image(&tx, &ty, &tz, box, tilt);
                } else {
                  tx = drx + -1. * Lx * rint(drx / Lx);
                  ty = dry + -1. * Ly * rint(dry / Ly);
                  tz = drz + -1. * Lz * rint(drz / Lz);
                }
                double dd = sqrt(tx * tx + ty * ty + tz * tz);
                if (dd == 0.0) continue;
                int64_t k = np_bin(dd, first, last, nbin, edges);
                if (k < 0) continue;
                int ta = type[i] < type[j] ? type[i] : type[j], tb = type[i] < type[j] ? type[j] : type[i];
                int64_t slot = (int64_t)ta * ntypes - (int64_t)ta * (ta - 1) / 2 + (tb - ta);
                h[slot * nbin + k]++;
              }
            }
      }
    }
#pragma omp critical
    for (int64_t k = 0; k < npair * nbin; ++k) hist[k] += h[k];
    free(h);
  }
  cells_free(&g);
  return 0;
}

int mo_rdf(const double *pos, const int32_t *type, const int64_t *mol, int64_t n, int ntypes,
           const double *box, double rcell, const double *edges, int64_t nbin, int same_molecule,
           int64_t *hist) {
  return mo_rdf_impl(pos, type, mol, n, ntypes, box, NULL, rcell, edges, nbin, same_molecule, hist);
}

int mo_rdf_tri(const double *pos, const int32_t *type, const int64_t *mol, int64_t n, int ntypes,
               const double *box, const double *tilt, double rcell, const double *edges, int64_t nbin,
               int same_molecule, int64_t *hist) {
  return mo_rdf_impl(pos, type, mol, n, ntypes, box, tilt, rcell, edges, nbin, same_molecule, hist);
}
