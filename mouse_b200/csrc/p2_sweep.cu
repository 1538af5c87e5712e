// K3: the pair sweep.  Replaces calculateAveCosSqForReference (ordering_functions.py:91-119)
// for all reference bonds of a frame at once.
//
// Two kernels:
//  * p2_sweep_fast_kernel - one warp per home cell.  The candidates of the 27 neighbour cells
//    are concatenated and held in REGISTERS (SLOTS x 32 lanes, cell-local FP32 coordinates
//    pre-shifted by the periodic cell offset), the home cell's reference bonds are looped
//    warp-uniformly, every lane tests its candidates with 3 FADD + FMUL + 2 FFMA + compare.
//    Sure hits (d2 < rc2*(1-g)) go straight to the FP64 cos^2; candidates inside the FP32
//    guard band are re-decided by the reference's exact FP64 expression (mouse_rsq_exact).
//    cos^2 = (u_i . u_j)^2 on FP64 unit vectors; |u_i . u_j| < 1e-9 falls back to the exact
//    un-fused raw dot so that the reference's "cos^2 == 0 is masked" rule (:113) is bit-exact.
//    Per-reference sums are reduced with warp shuffles in a fixed order (deterministic).
//  * p2_sweep_exact_kernel - one thread per reference, all FP64, literally the reference's
//    expression per candidate.  Valid for ANY grid (fewer than 3 cells on an axis, rcut < 0,
//    rmax == 0 -> half diagonal); also the on-device cross-check of the fast kernel.
#include "common.cuh"

struct SweepArgs {
  GridDev g;
  const float4 *P;
  const double4 *U;
  const int32_t *idx;
  const int32_t *cell_start;
  const int32_t *cell_sorted;
  const uint8_t *flag;
  const double *vec;  // SoA [3][n], original order
  const double *ctr;
  int n;
  double *out_sum;
  int32_t *out_cnt;
  int64_t *counters;
  const int64_t *row_ptr;  // lists only
  int32_t *nbr_idx;
  int32_t *fix_list;       // home cells queued for p2_fixup_kernel (length in counters[3])
  unsigned *redo_mask;     // [FAST_MAX_PASSES][n] per sorted slot: lanes to re-decide exactly (fast sweep -> fixup)
  int same_molecule;
  int no_cutoff;
};

// ------------------------------------------------------------------ rare exact path (fast kernel)
// The reference's decision for ONE pair of sorted slots (si = reference, sj = candidate): returns
// cos^2 exactly as ordering_functions.py:95-113 evaluates it, or 0.0 when the pair is not counted
// (out of range, or cos^2 == 0 which the reference masks).
__device__ __forceinline__ double exact_pair(const int32_t *__restrict__ idx, const double *__restrict__ ctr,
                                             const double *__restrict__ vec, size_t n, const double *box, double rc2,
                                             int si, int sj) {
  const int bi = idx[si], bj = idx[sj];
  const double rsq = mouse_rsq_exact(ctr[bi], ctr[n + bi], ctr[2 * n + bi], ctr[bj], ctr[n + bj], ctr[2 * n + bj], box);
  if (!(rsq <= rc2)) return 0.0;
  const double ax = vec[bi], ay = vec[n + bi], az = vec[2 * n + bi];
  const double bx = vec[bj], by = vec[n + bj], bz = vec[2 * n + bj];
  const double dot = mouse_dot_exact(ax, ay, az, bx, by, bz);
  const double nr = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
  const double nj = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
  return __ddiv_rn(__ddiv_rn(__dmul_rn(dot, dot), nr), nj);
}

// candidate t of the concatenated 27-cell list -> sorted slot (largest k < 27 with pref[k] <= t)
__device__ __forceinline__ int cand_slot(const int *pref, const int *start, int t) {
  int k = 0;
#pragma unroll
  for (int step = 16; step > 0; step >>= 1) {
    const int k2 = k + step;
    if (k2 < 27 && pref[k2] <= t) k = k2;
  }
  return start[k] + (t - pref[k]);
}

#define FAST_MAX_PASSES 4   // redo masks are kept for this many candidate passes per home cell
#ifndef FAST_TARGET_WARPS
#define FAST_TARGET_WARPS 16   // resident warps per SM the register allocation is sized for
#endif
#ifndef FAST_BU
#define FAST_BU 2             // hits per phase-B trip (independent FP64 chains per lane)
#endif
#ifndef FAST_RUN
#define FAST_RUN 1             // consecutive home cells per CTA visit (L1 reuse of shared neighbour cells)
#endif

// neighbour table of a home cell: per neighbour k (lane k < 27) first sorted slot, exclusive prefix of the
// counts, periodic cell shift.  Returns T (all lanes) and the home cell's own prefix.
__device__ __forceinline__ int neighbour_table(const SweepArgs &a, int home, int lane, int *s_start, int *s_pref,
                                               float (*s_off)[32], int &home_pref) {
  const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
  const int cx = home % nx, cy = (home / nx) % ny, cz = home / (nx * ny);
  int st = 0, cn = 0;
  float ox = 0.f, oy = 0.f, oz = 0.f;
  if (lane < 27) {
    const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
    int mx = cx + dx, my = cy + dy, mz = cz + dz;
    mx += (mx < 0) ? nx : 0;
    mx -= (mx >= nx) ? nx : 0;
    my += (my < 0) ? ny : 0;
    my -= (my >= ny) ? ny : 0;
    mz += (mz < 0) ? nz : 0;
    mz -= (mz >= nz) ? nz : 0;
    const int cid = (mz * ny + my) * nx + mx;
    st = a.cell_start[cid];
    cn = a.cell_start[cid + 1] - st;
    ox = (float)dx * a.g.cwf[0];
    oy = (float)dy * a.g.cwf[1];
    oz = (float)dz * a.g.cwf[2];
  }
  int incl = cn;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(MOUSE_FULL_MASK, incl, o);
    if (lane >= o) incl += y;
  }
  const int T = __shfl_sync(MOUSE_FULL_MASK, incl, 31);
  home_pref = __shfl_sync(MOUSE_FULL_MASK, incl - cn, 13);  // neighbour 13 = the home cell itself
  __syncwarp();
  s_start[lane] = st;
  s_pref[lane] = incl - cn;      // lanes >= 27 hold T
  if (s_off) {
    s_off[0][lane] = ox;
    s_off[1][lane] = oy;
    s_off[2][lane] = oz;
  }
  __syncwarp();
  return T;
}

// One warp per home cell (CTA = 1 warp; CTAs visit runs of FAST_RUN consecutive home cells).
//   load  : the candidates of the 27 neighbour cells are concatenated into t = 0..T-1; lane t % 32 keeps
//           candidate t's cell-local FP32 coordinates (shifted by the periodic cell offset) in REGISTERS
//           (slot t / 32) and its FP64 unit vector (+ molecule code) in SHARED memory.  The home cell's
//           reference bonds are staged in shared memory in chunks of 32 (lane l owns reference chunk0 + l).
//   per reference bond (warp-uniform loop, no global memory traffic, no calls):
//     A   : every lane tests its candidates in FP32 (3 FADD, FMUL, 2 FFMA, FADD) and shifts the sign of
//           d2 - rc2 into a hit mask (one funnel shift); min |d2 - rc2| is tracked for the guard band.
//     B   : lanes walk their hit bits (warp-uniform trip count, branch-free body):
//           cos^2 = (u_i . u_j)^2 in FP64 from shared memory.
//     flag: a lane that saw a guard-band candidate or a near-zero dot drops its contribution for this
//           reference and sets its bit in the reference's redo mask (decided exactly by p2_fixup_kernel).
//     sum : warp-shuffle reduction in a fixed order; the result stays in the owning lane's registers.
//   flush : one store per reference (sum, count, redo mask).
template <int SLOTS, bool EXCL_MOL, bool LISTS>
__global__ void __launch_bounds__(32, (FAST_TARGET_WARPS > 32 ? 32 : FAST_TARGET_WARPS))
    p2_sweep_fast_kernel(const SweepArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_start[32];
  __shared__ int s_pref[32];
  __shared__ float s_off[3][32];
  __shared__ float4 s_home[32];          // the chunk's reference bonds: cell-local xyz + molecule code
  __shared__ double s_hu[3][32];         //                             FP64 unit vectors
  const int lane = threadIdx.x;
  static_assert(SLOTS % 4 == 0, "slots come in groups of 4");
  constexpr int NC = SLOTS * 32;         // candidates per pass
  constexpr int NCZ = NC + 32;           // + one all-zero row for exhausted lanes
  constexpr int GROUPS = SLOTS / 4;
  constexpr int BU = FAST_BU;
  double *su = reinterpret_cast<double *>(smem_raw) + lane;      // [3][NC], lane column
  int *smol = reinterpret_cast<int *>(smem_raw + sizeof(double) * 3 * NCZ) + lane;   // [NCZ]
  su[NC] = 0.0;
  su[NCZ + NC] = 0.0;
  su[2 * NCZ + NC] = 0.0;
  smol[NC] = 0;
  const float4 *__restrict__ P = a.P;
  const double4 *__restrict__ U = a.U;
  const int32_t *__restrict__ cell_start = a.cell_start;
  // guard band [rc2_lo, rc2_hi] around rc^2: centre / half width in FP32
  const float rc2_c = 0.5f * (a.g.rc2_lo + a.g.rc2_hi);
  const float rc2_w = 0.5f * (a.g.rc2_hi - a.g.rc2_lo) * 1.0001f;
  // zero-length bonds in the neighbour set turn every mean into NaN and add to every count
  // (0/0 in ordering_functions.py:112)
  const long long nzero = LISTS ? 0 : a.counters[0];
#ifdef MOUSE_PHASE_TIMERS
  long long tm_load = 0, tm_a = 0, tm_b = 0, tm_e = 0, tm0;
#define TM_START() tm0 = clock64()
#define TM_ADD(x) do { long long t1_ = clock64(); x += t1_ - tm0; tm0 = t1_; } while (0)
#else
#define TM_START()
#define TM_ADD(x)
#endif

  // persistent warps fetch runs of FAST_RUN consecutive home cells from a global ticket counter
  // (dynamic load balance; consecutive cells share 18 of their 27 neighbour cells -> L1/L2 reuse)
  const int nruns = (a.g.ncell + FAST_RUN - 1) / FAST_RUN;
  for (;;) {
    int run = 0;
    if (lane == 0) run = (int)atomicAdd((unsigned long long *)&a.counters[2], 1ULL);
    run = __shfl_sync(MOUSE_FULL_MASK, run, 0);
    if (run >= nruns) break;
    for (int home = run * FAST_RUN; home < min(a.g.ncell, (run + 1) * FAST_RUN); ++home) {
      const int hs = cell_start[home], he = cell_start[home + 1];
      if (hs == he) continue;
      TM_START();
      int home_pref;
      const int T = neighbour_table(a, home, lane, s_start, s_pref, s_off, home_pref);
      unsigned home_any = (!LISTS && T > FAST_MAX_PASSES * NC) ? 1u : 0u;   // needs p2_fixup_kernel?

      // (candidate passes beyond FAST_MAX_PASSES are left entirely to p2_fixup_kernel)
      for (int base = 0, pass = 0; base < T && (LISTS || pass < FAST_MAX_PASSES); base += NC, ++pass) {
        // candidate coordinates as FP32 pairs (slots 2p, 2p+1): phase A runs on packed f32x2 instructions
        float2 xr2[SLOTS / 2], yr2[SLOTS / 2], zr2[SLOTS / 2];
        __syncwarp();
#pragma unroll
        for (int c = 0; c < SLOTS; ++c) {
          const int t = base + 32 * c + lane;
          float px = 1e30f, py = 0.f, pz = 0.f;
          if (t < T) {
            int k = 0;  // largest k < 27 with pref[k] <= t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
              const int k2 = k + step;
              if (k2 < 27 && s_pref[k2] <= t) k = k2;
            }
            const int j = s_start[k] + (t - s_pref[k]);
            const float4 p = __ldg(P + j);
            const double4 u = U[j];
            px = p.x + s_off[0][k];
            py = p.y + s_off[1][k];
            pz = p.z + s_off[2][k];
            su[c * 32] = u.x;
            su[NCZ + c * 32] = u.y;
            su[2 * NCZ + c * 32] = u.z;
            if (EXCL_MOL) smol[c * 32] = __float_as_int(p.w);
          } else if (EXCL_MOL) {
            smol[c * 32] = 0;
          }
          if (c & 1) {
            xr2[c >> 1].y = px;
            yr2[c >> 1].y = py;
            zr2[c >> 1].y = pz;
          } else {
            xr2[c >> 1].x = px;
            yr2[c >> 1].x = py;
            zr2[c >> 1].x = pz;
          }
        }
        const int nslots = min(SLOTS, (T - base + 31) >> 5);
        const int ngroups = (nslots + 3) >> 2;

        for (int chunk0 = hs; chunk0 < he; chunk0 += 32) {
          const int nchunk = min(32, he - chunk0);
          int my_idx = 0;
          bool my_ref = false;
          __syncwarp();
          if (lane < nchunk) {
            const int ic = chunk0 + lane;
            my_idx = a.idx[ic];
            my_ref = (a.flag[ic] & 1) != 0;
            s_home[lane] = __ldg(P + ic);
            const double4 u = U[ic];
            s_hu[0][lane] = u.x;
            s_hu[1][lane] = u.y;
            s_hu[2][lane] = u.z;
          }
          const unsigned refmask = __ballot_sync(MOUSE_FULL_MASK, my_ref);
          double racc = 0.0;      // result of reference chunk0 + lane
          int rcnt = 0;
          unsigned rredo = 0;     // lanes whose contribution to reference chunk0 + lane must be decided exactly
          __syncwarp();
          TM_ADD(tm_load);

          for (int il = 0; il < nchunk; ++il) {
            if (!((refmask >> il) & 1u)) continue;
            const float4 pi = s_home[il];
            const int moli = __float_as_int(pi.w);
            const double uix = s_hu[0][il], uiy = s_hu[1][il], uiz = s_hu[2][il];
            // the reference bond is itself a candidate (home cell = neighbour 13): slot / lane of it
            const int tself = home_pref + (chunk0 + il - hs) - base;
            // ---- A: FP32 filter; slot c ends up in bit c (groups are walked from the last slot down)
            unsigned mask = 0;
            float band = 1e30f;
            const float2 npx = make_float2(-pi.x, -pi.x), npy = make_float2(-pi.y, -pi.y), npz = make_float2(-pi.z, -pi.z);
            const float2 nrc = make_float2(-rc2_c, -rc2_c);
#pragma unroll
            for (int g = GROUPS - 1; g >= 0; --g) {
              if (g < ngroups) {
#pragma unroll
                for (int pp = 2 * g + 1; pp >= 2 * g; --pp) {
                  const float2 dx = __fadd2_rn(xr2[pp], npx), dy = __fadd2_rn(yr2[pp], npy), dz = __fadd2_rn(zr2[pp], npz);
                  const float2 t = __fadd2_rn(__ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx))), nrc);  // < 0: inside
                  mask = __funnelshift_l(__float_as_uint(t.y), mask, 1);   // slot 2pp+1
                  mask = __funnelshift_l(__float_as_uint(t.x), mask, 1);   // slot 2pp
                  band = fminf(band, fminf(fabsf(t.x), fabsf(t.y)));
                }
              }
            }
            if (tself >= 0 && tself < NC && lane == (tself & 31)) mask &= ~(1u << (tself >> 5));
            TM_ADD(tm_a);
            // ---- B: FP64 cos^2 of the hits
            double acc = 0.0;
            unsigned m = mask;
            unsigned nearz = 0xffffffffu;
            const int iters = __reduce_max_sync(MOUSE_FULL_MASK, __popc(mask));
            // BU hits per trip: BU independent load -> FMA chains per lane (ILP); branch-free bodies,
            // idle lanes read slot 0 and contribute d * 0
            double acc2[BU];
#pragma unroll
            for (int u = 0; u < BU; ++u) acc2[u] = 0.0;
            for (int it = 0; it < iters; it += BU) {
#pragma unroll
              for (int u = 0; u < BU; ++u) {
                // highest hit bit first; an exhausted lane (m == 0) clamps to the last slot (finite data,
                // contributes d * 0): no branch, one bit scan
                // highest hit bit first; an exhausted lane (m == 0) is sent to the all-zero row SLOTS of the
                // table (d == 0 exactly): no branch, no select, one bit scan
                const bool act = m != 0;
                const int c = (int)min((unsigned)(31 - __clz((int)m)), (unsigned)SLOTS);
                m &= ~(1u << c);
                const double *q = su + c * 32;
                const double d = fma(uix, q[0], fma(uiy, q[NCZ], uiz * q[2 * NCZ]));
                bool use = act;
                double w = d;
                if (EXCL_MOL) {
                  const bool same = smol[c * 32] == moli;
                  mask &= ~((act && same) ? (1u << c) : 0u);
                  use = act && !same;
                  w = same ? 0.0 : d;
                }
                acc2[u] = fma(d, w, acc2[u]);
                // smallest |d| (as its high word) over the used hits: |d| < 1e-9 must be decided exactly (:113)
                if (use) nearz = min(nearz, (unsigned)__double2hiint(d) & 0x7fffffffu);
              }
            }
#pragma unroll
            for (int u = 0; u < BU; ++u) acc += acc2[u];
            bool redo = (band <= rc2_w) || (nearz < 0x3e112e0bu);
            TM_ADD(tm_b);
            if (LISTS) {
              // pair lists are extracted with every decision made in place (not on the timed path)
              if (redo) {
                mask = 0;
                for (int c = 0; c < nslots; ++c) {
                  const int t = base + 32 * c + lane;
                  if (t >= T) break;
                  const int sj = cand_slot(s_pref, s_start, t);
                  if (sj == chunk0 + il) continue;
                  if (EXCL_MOL && __float_as_int(P[sj].w) == moli) continue;
                  if (exact_pair(a.idx, a.ctr, a.vec, (size_t)a.n, a.g.box, a.g.rc2, chunk0 + il, sj) != 0.0)
                    mask |= 1u << c;
                }
              }
              __syncwarp();
              int cntl = __popc(mask);
              const int o = __shfl_sync(MOUSE_FULL_MASK, my_idx, il);
              int64_t row = a.row_ptr[o] + a.out_cnt[o];  // out_cnt = entries written by earlier passes
              int incl_c = cntl;
#pragma unroll
              for (int off = 1; off < 32; off <<= 1) {
                int y = __shfl_up_sync(MOUSE_FULL_MASK, incl_c, off);
                if (lane >= off) incl_c += y;
              }
              row += incl_c - cntl;
              unsigned mm = mask;
              while (mm) {
                const int c = __ffs(mm) - 1;
                mm &= mm - 1;
                a.nbr_idx[row++] = a.idx[cand_slot(s_pref, s_start, base + 32 * c + lane)];
              }
              cntl = __reduce_add_sync(MOUSE_FULL_MASK, cntl);
              if (lane == 0) a.out_cnt[o] += cntl;
              __syncwarp();
              continue;
            }
            // ---- flag: flagged lanes drop their contribution; p2_fixup_kernel adds the exact one
            const unsigned fm = __ballot_sync(MOUSE_FULL_MASK, redo);
            home_any |= fm;
            if (redo) {
              acc = 0.0;
              mask = 0;
            }
            acc = warp_sum(acc);
            const int cnt = __reduce_add_sync(MOUSE_FULL_MASK, __popc(mask));
            if (lane == il) {
              racc = acc;
              rcnt = cnt;
              rredo = fm;
            }
            TM_ADD(tm_e);
          }
          // flush the chunk: one store per reference, no global read unless a second pass is running
          if (!LISTS && lane < nchunk) {
            if (my_ref) {
              if (base == 0) {
                if (nzero) {
                  racc = __longlong_as_double(0x7ff8000000000000LL);
                  rcnt += (int)nzero;
                }
                a.out_sum[my_idx] = racc;
                a.out_cnt[my_idx] = rcnt;
              } else {
                a.out_sum[my_idx] += racc;
                a.out_cnt[my_idx] += rcnt;
              }
            }
            a.redo_mask[(size_t)pass * a.n + chunk0 + lane] = rredo;
          }
          TM_ADD(tm_e);
        }
      }
      // home cells with dropped contributions (or skipped passes) are queued for p2_fixup_kernel
      if (!LISTS && home_any && lane == 0)
        a.fix_list[atomicAdd((unsigned long long *)&a.counters[3], 1ULL)] = home;
    }
  }
#ifdef MOUSE_PHASE_TIMERS
  if (lane == 0) {
    atomicAdd((unsigned long long *)&a.counters[4], (unsigned long long)tm_load);
    atomicAdd((unsigned long long *)&a.counters[5], (unsigned long long)tm_a);
    atomicAdd((unsigned long long *)&a.counters[6], (unsigned long long)tm_b);
    atomicAdd((unsigned long long *)&a.counters[7], (unsigned long long)tm_e);
  }
#endif
}

// Exact re-decision of the contributions the fast sweep dropped: one warp per home cell; for every reference
// with a non-zero redo mask, every flagged lane's candidates (slots 0..SLOTS-1 of that pass) are evaluated with
// the reference's FP64 expressions - lane s takes slot s - and added to the reference's sum / count in a fixed
// order (pass, lane, slot tree), so the result stays deterministic.  Candidate passes beyond FAST_MAX_PASSES
// (more than FAST_MAX_PASSES * SLOTS * 32 candidates around one cell) are recomputed exactly as a whole.
template <int SLOTS, bool EXCL_MOL>
__global__ void __launch_bounds__(32) p2_fixup_kernel(const SweepArgs a) {
  __shared__ int s_start[32];
  __shared__ int s_pref[32];
  const int lane = threadIdx.x;
  constexpr int NC = SLOTS * 32;
  const size_t n = (size_t)a.n;
  const int nfix = (int)((volatile int64_t *)a.counters)[3];
  for (int e = blockIdx.x; e < nfix; e += gridDim.x) {
    const int home = a.fix_list[e];
    const int hs = a.cell_start[home], he = a.cell_start[home + 1];
    int home_pref = 0;
    const int T = neighbour_table(a, home, lane, s_start, s_pref, nullptr, home_pref);
    const int npass = (T + NC - 1) / NC;
    for (int chunk0 = hs; chunk0 < he; chunk0 += 32) {
      // lane l reads the redo masks of reference chunk0 + l (coalesced), flagged references are then
      // visited one at a time
      const int i_l = chunk0 + lane;
      unsigned mk[FAST_MAX_PASSES];
      bool want = false;
#pragma unroll
      for (int pass = 0; pass < FAST_MAX_PASSES; ++pass) {
        mk[pass] = (i_l < he && pass < npass) ? a.redo_mask[(size_t)pass * n + i_l] : 0u;
        want |= mk[pass] != 0;
      }
      want = (i_l < he) && (a.flag[min(i_l, he - 1)] & 1) && (want || npass > FAST_MAX_PASSES);
      unsigned todo = __ballot_sync(MOUSE_FULL_MASK, want);
      while (todo) {
        const int il = __ffs(todo) - 1;
        todo &= todo - 1;
        const int i = chunk0 + il;
        const int moli = __float_as_int(a.P[i].w);
        double add = 0.0;
        int addc = 0;
        for (int pass = 0; pass < npass; ++pass) {
          // passes the fast sweep skipped are evaluated for every lane
          unsigned lanes = 0xffffffffu;
#pragma unroll
          for (int q = 0; q < FAST_MAX_PASSES; ++q)
            if (q == pass) lanes = __shfl_sync(MOUSE_FULL_MASK, mk[q], il);
          while (lanes) {
            const int L = __ffs(lanes) - 1;
            lanes &= lanes - 1;
            const int t = pass * NC + 32 * lane + L;          // lane s handles slot s of flagged lane L
            double cs = 0.0;
            if (lane < SLOTS && t < T) {
              const int sj = cand_slot(s_pref, s_start, t);
              if (sj != i && !(EXCL_MOL && __float_as_int(a.P[sj].w) == moli))
                cs = exact_pair(a.idx, a.ctr, a.vec, n, a.g.box, a.g.rc2, i, sj);
            }
            add += warp_sum(cs);
            addc += __reduce_add_sync(MOUSE_FULL_MASK, cs != 0.0 ? 1 : 0);
            if (lane == 0) atomicAdd((unsigned long long *)&a.counters[1], 1ULL);
          }
        }
        if (lane == 0 && (add != 0.0 || addc != 0)) {
          const int o = a.idx[i];
          a.out_sum[o] += add;
          a.out_cnt[o] += addc;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ exact all-FP64 sweep
template <bool LISTS>
__global__ void __launch_bounds__(128) p2_sweep_exact_kernel(const SweepArgs a) {
  const int n_list = a.cell_start[a.g.ncell];
  const size_t n = (size_t)a.n;
  const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_list; s += gridDim.x * blockDim.x) {
    if (!(a.flag[s] & 1)) continue;
    const int bi = a.idx[s];
    const double cix = a.ctr[bi], ciy = a.ctr[n + bi], ciz = a.ctr[2 * n + bi];
    const double bix = a.vec[bi], biy = a.vec[n + bi], biz = a.vec[2 * n + bi];
    const double nr = __dadd_rn(__dadd_rn(__dmul_rn(bix, bix), __dmul_rn(biy, biy)), __dmul_rn(biz, biz));
    const int moli = __float_as_int(a.P[s].w);
    const int home = a.cell_sorted[s];
    const int cx = home % nx, cy = (home / nx) % ny, cz = home / (nx * ny);
    // neighbour cells per axis: {c-1, c, c+1} wrapped when the axis has >= 3 cells, else all cells
    const int mx = nx >= 3 ? 3 : nx, my = ny >= 3 ? 3 : ny, mz = nz >= 3 ? 3 : nz;
    double sum = 0.0;
    int cnt = 0;
    int64_t row = 0;
    if (LISTS) row = a.row_ptr[bi];
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
            if (!a.same_molecule && __float_as_int(a.P[t].w) == moli) continue;
            const int bj = a.idx[t];
            if (!a.no_cutoff) {
              double rsq = mouse_rsq_exact(cix, ciy, ciz, a.ctr[bj], a.ctr[n + bj], a.ctr[2 * n + bj], a.g.box);
              if (!(rsq <= a.g.rc2)) continue;
            }
            const double dot = mouse_dot_exact(bix, biy, biz, a.vec[bj], a.vec[n + bj], a.vec[2 * n + bj]);
            const double nj = a.U[t].w;
            const double cs = __ddiv_rn(__ddiv_rn(__dmul_rn(dot, dot), nr), nj);
            if (cs != 0.0) {
              if (LISTS) a.nbr_idx[row + cnt] = bj;
              sum += cs;
              cnt += 1;
            }
          }
        }
      }
    }
    if (!LISTS) {
      const long long nzero = a.counters[0];
      if (nzero) {
        sum = __longlong_as_double(0x7ff8000000000000LL);
        cnt += (int)nzero;
      }
      a.out_sum[bi] = sum;
      a.out_cnt[bi] = cnt;
    }
  }
}

// ------------------------------------------------------------------ row_ptr = exclusive scan of cnt (lists)
__global__ void __launch_bounds__(1024) row_ptr_kernel(const int32_t *__restrict__ cnt, int n,
                                                       int64_t *__restrict__ row_ptr) {
  // single block, sequential chunks; only used on the (non-hot) pair-list extraction path
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    int k = base + threadIdx.x;
    long long v = k < n ? cnt[k] : 0;
    long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long y = __shfl_up_sync(MOUSE_FULL_MASK, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      long long w = s_warp[threadIdx.x], xs = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long y = __shfl_up_sync(MOUSE_FULL_MASK, xs, o);
        if (threadIdx.x >= o) xs += y;
      }
      s_warp[threadIdx.x] = xs - w;
    }
    __syncthreads();
    long long carry = s_carry;
    long long incl = x + s_warp[threadIdx.x >> 5];
    if (k < n) row_ptr[k] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) row_ptr[n] = s_carry;
}

// ------------------------------------------------------------------ launchers
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int mouse_fast_variant = 1;  // tuning / debug knob (2 = skip the fixup kernel)

template <int SLOTS, bool EXCL, bool LISTS>
static void launch_fast_one(const SweepArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)(SLOTS + 1) * 32 * (3 * sizeof(double) + sizeof(int));
  static bool configured = false;
  static int blocks_per_sm = 1;
  if (!configured) {
    cudaFuncSetAttribute(p2_sweep_fast_kernel<SLOTS, EXCL, LISTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, p2_sweep_fast_kernel<SLOTS, EXCL, LISTS>, 32, smem);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  // persistent-style grid: a few waves of resident warps, each striding over runs of home cells
  const int nruns = (a.g.ncell + FAST_RUN - 1) / FAST_RUN;
  int blocks = num_sms() * blocks_per_sm;
  if (blocks > nruns) blocks = nruns;
  p2_sweep_fast_kernel<SLOTS, EXCL, LISTS><<<blocks, 32, smem, st>>>(a);
  if (!LISTS && mouse_fast_variant != 2) {   // variant 2 (debug): leave the dropped contributions out
    int fb = num_sms() * 32;
    if (fb > a.g.ncell) fb = a.g.ncell;
    p2_fixup_kernel<SLOTS, EXCL><<<fb, 32, 0, st>>>(a);
  }
}


template <int SLOTS>
static void launch_fast(const SweepArgs &a, bool excl_mol, bool lists, cudaStream_t st) {
  if (lists) {
    if (excl_mol) launch_fast_one<SLOTS, true, true>(a, st);
    else launch_fast_one<SLOTS, false, true>(a, st);
  } else {
    if (excl_mol) launch_fast_one<SLOTS, true, false>(a, st);
    else launch_fast_one<SLOTS, false, false>(a, st);
  }
}

int mouse_fast_slots_override = 0;  // test / tuning hooks (set through mouse_set_tuning)

int mouse_launch_p2_sweep(const double *vec, const double *ctr, int64_t n, const mouse_grid_t *grid, uint32_t flags,
                          const Workspace &w, double *out_sum, int32_t *out_cnt, const int64_t *row_ptr,
                          int32_t *nbr_idx, cudaStream_t st) {
  SweepArgs a;
  a.g = mouse_grid_dev(grid);
  a.P = w.p_sorted;
  a.U = w.u_sorted;
  a.idx = w.idx_sorted;
  a.cell_start = w.cell_start;
  a.cell_sorted = w.cell_sorted;
  a.flag = w.flag_sorted;
  a.vec = vec;
  a.ctr = ctr;
  a.n = (int)n;
  a.out_sum = out_sum;
  a.out_cnt = out_cnt;
  a.counters = w.counters;
  a.row_ptr = row_ptr;
  a.nbr_idx = nbr_idx;
  a.fix_list = w.cell_count;           // cell_count is free once cell_start exists
  a.redo_mask = (unsigned *)w.p_tmp;   // p_tmp (16 B per item) is free after the cell build: 4 x u32 per item
  a.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
  a.no_cutoff = (flags & MOUSE_NO_CUTOFF) ? 1 : 0;
  const bool lists = nbr_idx != nullptr;
  if (n == 0) return MOUSE_OK;
  if (!lists) {
    MOUSE_CUDA_CHECK(cudaMemsetAsync(out_sum, 0, sizeof(double) * (size_t)n, st));
  }
  MOUSE_CUDA_CHECK(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t) * (size_t)n, st));
  // counters[1] = exact re-decisions (diagnostic), counters[3] = length of the fixup list: per sweep
  // counters[2] = run ticket of the persistent fast sweep (the reduce kernel resets it before using it)
  MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters + 1, 0, 3 * sizeof(int64_t), st));
  const bool fast = grid->fast && !(flags & (MOUSE_FORCE_EXACT | MOUSE_NO_CUTOFF));
  if (fast) {
    // SLOTS*32 registers-resident candidates per pass: size for mean + ~3 sigma of 27 cells
    double mean = 27.0 * (double)n / (double)grid->ncell;
    int need = (int)((1.15 * mean + 2.0 * sqrt(mean) + 31.0) / 32.0);
    int slots = need <= 4 ? 4 : need <= 8 ? 8 : need <= 12 ? 12 : 16;
    (void)need;
    if (mouse_fast_slots_override) slots = mouse_fast_slots_override;
    const bool excl = !a.same_molecule;
    switch (slots) {
      case 4: launch_fast<4>(a, excl, lists, st); break;
      case 8: launch_fast<8>(a, excl, lists, st); break;
      case 12: launch_fast<12>(a, excl, lists, st); break;
      default: launch_fast<16>(a, excl, lists, st); break;
    }
    MOUSE_LAUNCH_CHECK("p2_sweep_fast_kernel");
  } else {
    int blocks = (int)((n + 127) / 128);
    int maxb = num_sms() * 32;
    if (blocks > maxb) blocks = maxb;
    if (lists)
      p2_sweep_exact_kernel<true><<<blocks, 128, 0, st>>>(a);
    else
      p2_sweep_exact_kernel<false><<<blocks, 128, 0, st>>>(a);
    MOUSE_LAUNCH_CHECK("p2_sweep_exact_kernel");
  }
  return MOUSE_OK;
}

int mouse_launch_row_ptr(const int32_t *cnt, int64_t n, int64_t *row_ptr, cudaStream_t st) {
  row_ptr_kernel<<<1, 1024, 0, st>>>(cnt, (int)n, row_ptr);
  MOUSE_LAUNCH_CHECK("row_ptr_kernel");
  return MOUSE_OK;
}
