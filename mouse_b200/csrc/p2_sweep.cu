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
  int same_molecule;
  int no_cutoff;
};

// ------------------------------------------------------------------ rare exact path (fast kernel)
// Re-decides ALL candidates of one lane with the reference's FP64 expressions.  Called only for
// lanes that saw a candidate inside the FP32 guard band around rc^2 or a near-zero dot product.
// Returns the exact hit mask over the lane's slots (slot c -> bit c) and the exact sum of cos^2.
struct RedoOut {
  unsigned mask;
  double acc;
};

__device__ __noinline__ RedoOut exact_redo(const int32_t *__restrict__ idx, const double *__restrict__ ctr,
                                           const double *__restrict__ vec, const float4 *__restrict__ P, size_t n,
                                           const double *box, double rc2, int si, const int *pref, const int *start,
                                           int tbase, int tstride, int nslots, bool excl_mol,
                                           unsigned long long *band_counter) {
  const int bi = idx[si];
  const double cix = ctr[bi], ciy = ctr[n + bi], ciz = ctr[2 * n + bi];
  const double ax = vec[bi], ay = vec[n + bi], az = vec[2 * n + bi];
  const double nr = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
  const int moli = __float_as_int(P[si].w);
  const int T = pref[27];
  RedoOut out;
  out.mask = 0;
  out.acc = 0.0;
  for (int c = 0; c < nslots; ++c) {
    const int t = tbase + tstride * c;   // this lane's candidate in slot c
    if (t >= T) break;
    int k = 0;
    for (int step = 16; step > 0; step >>= 1) {
      const int k2 = k + step;
      if (k2 < 27 && pref[k2] <= t) k = k2;
    }
    const int sjv = start[k] + (t - pref[k]);
    if (sjv == si) continue;
    if (excl_mol && __float_as_int(P[sjv].w) == moli) continue;
    const int bj = idx[sjv];
    const double rsq = mouse_rsq_exact(cix, ciy, ciz, ctr[bj], ctr[n + bj], ctr[2 * n + bj], box);
    if (!(rsq <= rc2)) continue;
    const double bx = vec[bj], by = vec[n + bj], bz = vec[2 * n + bj];
    const double dot = mouse_dot_exact(ax, ay, az, bx, by, bz);
    const double nj = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
    const double cs = __ddiv_rn(__ddiv_rn(__dmul_rn(dot, dot), nr), nj);
    if (cs != 0.0) {
      out.mask |= 1u << c;
      out.acc += cs;
    }
  }
  atomicAdd(band_counter, 1ULL);
  return out;
}

#ifndef FAST_TARGET_WARPS
#define FAST_TARGET_WARPS 16   // resident warps per SM the register allocation is sized for
#endif

// One CTA of K warps per home cell (grid-stride over home cells).
//   load  : the candidates of the 27 neighbour cells are concatenated into t = 0..T-1 and dealt out in
//           chunks of 32: chunk g belongs to warp g % K, slot g / K, lane t % 32.  Each lane keeps its
//           candidates' cell-local FP32 coordinates (shifted by the periodic cell offset) in REGISTERS and
//           their FP64 unit vectors (+ molecule codes) in SHARED memory.  The home cell's reference bonds
//           are staged in shared memory in chunks of 32 (lane l of every warp owns reference chunk0 + l).
//   per reference bond (warp-uniform loop, no global memory traffic):
//     A   : every lane tests its candidates in FP32 (3 FADD, FMUL, 2 FFMA, FADD) and shifts the sign of
//           d2 - rc2 into a hit mask (one funnel shift); min |d2 - rc2| is tracked for the guard band.
//     B   : lanes walk their hit bits (warp-uniform trip count, branch-free body):
//           cos^2 = (u_i . u_j)^2 in FP64 from shared memory.
//     fix : lanes that saw a guard-band candidate or a near-zero dot redo their slots exactly.
//     sum : warp-shuffle reduction in a fixed order; the result stays in the owning lane's registers.
//   flush : the K warps' partial results are combined in warp order through shared memory (deterministic)
//           and stored with one write per reference.
template <int SLOTS, int K, bool EXCL_MOL, bool LISTS>
__global__ void __launch_bounds__(K * 32, (FAST_TARGET_WARPS / K > 32 ? 32 : FAST_TARGET_WARPS / K)) p2_sweep_fast_kernel(const SweepArgs a) {
  static_assert(!LISTS || K == 1, "pair lists use the single-warp variant");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_start[32];
  __shared__ int s_pref[32];
  __shared__ float s_off[3][32];
  __shared__ float4 s_home[32];          // the chunk's reference bonds: cell-local xyz + molecule code
  __shared__ double s_hu[3][32];         //                             FP64 unit vectors
  __shared__ double s_pacc[K][32];       // partial results of warps 1..K-1
  __shared__ int s_pcnt[K][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  constexpr int NC = SLOTS * 32;         // candidates per warp and pass
  constexpr int GROUPS = (SLOTS + 3) / 4;
  double *su = reinterpret_cast<double *>(smem_raw) + (size_t)wib * (3 * NC) + lane;      // [3][NC], lane column
  int *smol = reinterpret_cast<int *>(smem_raw + sizeof(double) * 3 * NC * K) + wib * NC + lane;
  const float4 *__restrict__ P = a.P;
  const double4 *__restrict__ U = a.U;
  const int32_t *__restrict__ cell_start = a.cell_start;
  const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
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

  for (int home = blockIdx.x; home < a.g.ncell; home += gridDim.x) {
    const int hs = cell_start[home], he = cell_start[home + 1];
    if (hs == he) continue;          // CTA-uniform
    TM_START();
    const int cx = home % nx, cy = (home / nx) % ny, cz = home / (nx * ny);
    // neighbour table (every warp computes it redundantly in registers; warp 0 publishes it)
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
      st = cell_start[cid];
      cn = cell_start[cid + 1] - st;
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
    const int home_pref = __shfl_sync(MOUSE_FULL_MASK, incl - cn, 13);  // neighbour 13 = the home cell itself
    __syncthreads();                 // previous home cell fully consumed
    if (wib == 0) {
      s_start[lane] = st;
      s_pref[lane] = incl - cn;      // lanes >= 27 hold T (sentinel, read by exact_redo as pref[27])
      s_off[0][lane] = ox;
      s_off[1][lane] = oy;
      s_off[2][lane] = oz;
    }
    __syncthreads();

    for (int base = 0; base < T; base += NC * K) {
      float xr[SLOTS], yr[SLOTS], zr[SLOTS];
#pragma unroll
      for (int c = 0; c < SLOTS; ++c) {
        const int t = base + 32 * (c * K + wib) + lane;
        xr[c] = 1e30f;
        yr[c] = 0.f;
        zr[c] = 0.f;
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
          xr[c] = p.x + s_off[0][k];
          yr[c] = p.y + s_off[1][k];
          zr[c] = p.z + s_off[2][k];
          su[c * 32] = u.x;
          su[NC + c * 32] = u.y;
          su[2 * NC + c * 32] = u.z;
          if (EXCL_MOL) smol[c * 32] = __float_as_int(p.w);
        }
      }
      // slots of this warp that hold candidates in this pass
      const int nchunks = min(SLOTS * K, (T - base + 31) >> 5);
      const int nslots = (nchunks - wib + K - 1) / K;
      const int ngroups = (nslots + 3) >> 2;

      for (int chunk0 = hs; chunk0 < he; chunk0 += 32) {
        const int nchunk = min(32, he - chunk0);
        int my_idx = 0;
        bool my_ref = false;
        __syncthreads();             // staging buffers free (and, first time, su complete per warp)
        if (lane < nchunk) {
          const int ic = chunk0 + lane;
          my_idx = a.idx[ic];
          my_ref = (a.flag[ic] & 1) != 0;
          if (wib == 0) {
            s_home[lane] = __ldg(P + ic);
            const double4 u = U[ic];
            s_hu[0][lane] = u.x;
            s_hu[1][lane] = u.y;
            s_hu[2][lane] = u.z;
          }
        }
        const unsigned refmask = __ballot_sync(MOUSE_FULL_MASK, my_ref);
        double racc = 0.0;   // this warp's partial result of reference chunk0 + lane
        int rcnt = 0;
        __syncthreads();
        TM_ADD(tm_load);

        for (int il = 0; il < nchunk; ++il) {
          if (!((refmask >> il) & 1u)) continue;
          const float4 pi = s_home[il];
          const int moli = __float_as_int(pi.w);
          const double uix = s_hu[0][il], uiy = s_hu[1][il], uiz = s_hu[2][il];
          // the reference bond is itself a candidate (home cell = neighbour 13): chunk / lane of its slot
          const int tself = home_pref + (chunk0 + il - hs) - base;
          // ---- A: FP32 filter; slot c ends up in bit c (groups are walked from the last slot down)
          unsigned mask = 0;
          float band = 1e30f;
#pragma unroll
          for (int g = GROUPS - 1; g >= 0; --g) {
            if (g < ngroups) {
#pragma unroll
              for (int c = min(4 * g + 3, SLOTS - 1); c >= 4 * g; --c) {
                const float dx = xr[c] - pi.x, dy = yr[c] - pi.y, dz = zr[c] - pi.z;
                const float t = fmaf(dz, dz, fmaf(dy, dy, dx * dx)) - rc2_c;   // < 0: inside
                mask = __funnelshift_l(__float_as_uint(t), mask, 1);
                band = fminf(band, fabsf(t));
              }
            }
          }
          if (SLOTS % 4 != 0) mask >>= 0;   // (slot numbering is unaffected by the partial last group)
          {
            const int gself = tself >> 5;   // chunk that holds the reference itself
            if (tself >= 0 && gself < SLOTS * K && (gself % K) == wib && lane == (tself & 31))
              mask &= ~(1u << (gself / K));
          }
          TM_ADD(tm_a);
          // ---- B: FP64 cos^2 of the hits
          bool redo = band <= rc2_w;
          double acc = 0.0;
          unsigned m = mask;
          unsigned nearz = 0xffffffffu;
          const int iters = __reduce_max_sync(MOUSE_FULL_MASK, __popc(mask));
          for (int it = 0; it < iters; ++it) {
            // branch-free body: idle lanes read slot 0 and contribute d * 0
            const bool act = m != 0;
            const int c = act ? (__ffs(m) - 1) : 0;
            m &= m - 1;
            const double *q = su + c * 32;
            const double d = fma(uix, q[0], fma(uiy, q[NC], uiz * q[2 * NC]));
            bool use = act;
            if (EXCL_MOL) {
              const bool same = smol[c * 32] == moli;
              mask &= ~((act && same) ? (1u << c) : 0u);
              use = act && !same;
            }
            const double w = use ? d : 0.0;
            acc = fma(d, w, acc);
            // smallest |d| (as its high word) over the used hits: |d| < 1e-9 must be decided exactly (:113)
            const unsigned h = use ? ((unsigned)__double2hiint(d) & 0x7fffffffu) : 0xffffffffu;
            nearz = min(nearz, h);
          }
          redo |= nearz < 0x3e112e0bu;
          TM_ADD(tm_b);
          // ---- fix: exact re-decision for the (rare) flagged lanes
          if (__any_sync(MOUSE_FULL_MASK, redo)) {
            if (redo) {
              const RedoOut ro = exact_redo(a.idx, a.ctr, a.vec, P, (size_t)a.n, a.g.box, a.g.rc2, chunk0 + il, s_pref,
                                            s_start, base + 32 * wib + lane, 32 * K, nslots, EXCL_MOL,
                                            (unsigned long long *)&a.counters[1]);
              mask = ro.mask;
              acc = ro.acc;
            }
          }
          int cnt = __popc(mask);
          if (LISTS) {
            const int o = __shfl_sync(MOUSE_FULL_MASK, my_idx, il);
            int64_t row = a.row_ptr[o] + a.out_cnt[o];  // out_cnt = entries written by earlier passes
            int incl_c = cnt;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              int y = __shfl_up_sync(MOUSE_FULL_MASK, incl_c, off);
              if (lane >= off) incl_c += y;
            }
            row += incl_c - cnt;
            unsigned mm = mask;
            while (mm) {
              const int c = __ffs(mm) - 1;
              mm &= mm - 1;
              const int t = base + 32 * (c * K + wib) + lane;
              int k = 0;
              for (int step = 16; step > 0; step >>= 1) {
                const int k2 = k + step;
                if (k2 < 27 && s_pref[k2] <= t) k = k2;
              }
              a.nbr_idx[row++] = a.idx[s_start[k] + (t - s_pref[k])];
            }
            __syncwarp();
          }
          acc = warp_sum(acc);
          cnt = __reduce_add_sync(MOUSE_FULL_MASK, cnt);
          if (lane == il) {
            racc = acc;
            rcnt = cnt;
          }
          TM_ADD(tm_e);
        }
        // flush the chunk: combine the K warps' partials in warp order, one store per reference
        if (K > 1) {
          if (wib > 0) {
            s_pacc[wib][lane] = racc;
            s_pcnt[wib][lane] = rcnt;
          }
          __syncthreads();
          if (wib == 0) {
#pragma unroll
            for (int w = 1; w < K; ++w) {
              racc += s_pacc[w][lane];
              rcnt += s_pcnt[w][lane];
            }
          }
        }
        if (wib == 0 && my_ref) {
          if (LISTS) {
            a.out_cnt[my_idx] += rcnt;
          } else if (base == 0) {
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
        TM_ADD(tm_e);
      }
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

template <int SLOTS, int K, bool EXCL, bool LISTS>
static void launch_fast_one(const SweepArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)K * SLOTS * 32 * (3 * sizeof(double) + sizeof(int));
  static bool configured = false;
  static int blocks_per_sm = 1;
  if (!configured) {
    cudaFuncSetAttribute(p2_sweep_fast_kernel<SLOTS, K, EXCL, LISTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, p2_sweep_fast_kernel<SLOTS, K, EXCL, LISTS>, K * 32,
                                                  smem);
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    configured = true;
  }
  // persistent-style grid: a few waves of resident CTAs, each striding over home cells
  int blocks = num_sms() * blocks_per_sm * 4;
  if (blocks > a.g.ncell) blocks = a.g.ncell;
  p2_sweep_fast_kernel<SLOTS, K, EXCL, LISTS><<<blocks, K * 32, smem, st>>>(a);
}

int mouse_fast_variant = 1;  // K = warps per home cell (1, 2 or 4); 1 measured fastest on B200

// total candidate slots per lane over the CTA = SLOTS * K
template <int TOTAL>
static void launch_fast(const SweepArgs &a, bool excl_mol, bool lists, cudaStream_t st) {
  if (lists) {
    if (excl_mol) launch_fast_one<TOTAL, 1, true, true>(a, st);
    else launch_fast_one<TOTAL, 1, false, true>(a, st);
    return;
  }
  const int k = mouse_fast_variant;
  if (k >= 4) {
    if (excl_mol) launch_fast_one<TOTAL / 4, 4, true, false>(a, st);
    else launch_fast_one<TOTAL / 4, 4, false, false>(a, st);
  } else if (k >= 2) {
    if (excl_mol) launch_fast_one<TOTAL / 2, 2, true, false>(a, st);
    else launch_fast_one<TOTAL / 2, 2, false, false>(a, st);
  } else {
    if (excl_mol) launch_fast_one<TOTAL, 1, true, false>(a, st);
    else launch_fast_one<TOTAL, 1, false, false>(a, st);
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
  a.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
  a.no_cutoff = (flags & MOUSE_NO_CUTOFF) ? 1 : 0;
  const bool lists = nbr_idx != nullptr;
  if (n == 0) return MOUSE_OK;
  if (!lists) {
    MOUSE_CUDA_CHECK(cudaMemsetAsync(out_sum, 0, sizeof(double) * (size_t)n, st));
  }
  MOUSE_CUDA_CHECK(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t) * (size_t)n, st));
  const bool fast = grid->fast && !(flags & (MOUSE_FORCE_EXACT | MOUSE_NO_CUTOFF));
  if (fast) {
    // SLOTS*32 registers-resident candidates per pass: size for mean + ~3 sigma of 27 cells
    double mean = 27.0 * (double)n / (double)grid->ncell;
    int need = (int)((mean + 1.0 * sqrt(mean) + 31.0) / 32.0);
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
