// K5: RDF pair histogram on the same cell list.  Replaces calculateRdfForReference
// (ordering_functions.py:66-88) summed over all reference atoms as in
// CalculatePairCorrelationFunctions (:162-173), counts accumulated onto zeros.
//
// v1: one thread per reference atom, all-FP64 distance in the reference's operation order
// (np.add(x, -1.*x_ref); dr + -1.*L*around(dr/L); sqrt(x^2+y^2+z^2)), NumPy's uniform-bin index
// rule (truncate, then +-1 correction against the edges table), shared-memory privatised
// histogram per block flushed with integer atomics (deterministic result).
#include "half_shell.cuh"

struct RdfArgs {
  GridDev g;
  const Rec *R;          // .p.w = type code; .u = absolute position (FP64) + molecule code in .w
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
    const double4 pi = a.R[s].u;
    const int ti = __float_as_int(a.R[s].p.w);
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
            const double4 pj = a.R[t].u;
            if (!a.same_molecule && pj.w == pi.w) continue;
            double tx = __dsub_rn(pj.x, pi.x), ty = __dsub_rn(pj.y, pi.y), tz = __dsub_rn(pj.z, pi.z);
            if (a.g.tri) {
              mouse_min_image_tri(tx, ty, tz, a.g.box, a.g.tilt);
            } else {
              tx = mouse_trim(tx, Lx);
              ty = mouse_trim(ty, Ly);
              tz = mouse_trim(tz, Lz);
            }
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz)));
            if (d == 0.0) continue;  // np.ma.masked_equal(d, 0.) also drops coincident atoms (:84)
            const int k = np_bin(d, first, last, a.nbin, a.edges);
            if (k < 0) continue;
            const int tj = __float_as_int(a.R[t].p.w);
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

// ------------------------------------------------------------------ fast RDF sweep (half shell)
// Same decomposition as the P2 sweep (half_shell.cuh): one warp per home cell, every unordered pair of atoms is
// decided once and counted twice (the reference counts both orderings into the same sorted type-pair key,
// ordering_functions.py:162-173).  Candidates come through TMA into shared memory; their FP32 coordinates are held
// in registers for a packed f32x2 distance filter against a slightly inflated rmax^2; the lanes then walk their
// filter hits and evaluate the reference's FP64 expression exactly (un-fused dr, minimum image, (x^2+y^2)+z^2).
// The bin is found WITHOUT sqrt / division in FP64: thr[k] is the smallest double s whose correctly rounded
// sqrt(s) is >= edges[k] (rdf_thresholds_kernel), so "edges[k] <= sqrt(s) < edges[k+1]" is decided by comparing s
// itself - bit-exact with numpy.histogram's rule, including range ends and the closed last bin.
#define RH_WARPS 16        // resident warps per SM the register allocation is sized for
#ifndef RH_CTA_WARPS
#define RH_CTA_WARPS 2      // warps per CTA: they share one privatised histogram
#endif
#ifndef RH_SLOTS
#define RH_SLOTS 6          // register-resident candidate slots per lane (candidates per landing-zone pass = 32 * RH_SLOTS):
                            // 6 -> 12 warps per SM (4M beads: 11.6 ms; 8 slots, 11 warps: 11.9 ms; 4 slots: 12.2 ms)
#endif
#ifndef RH_WALK
#define RH_WALK 2          // queued hits per lane and walk round
#endif
#ifndef RH_HOMEBUF
#define RH_HOMEBUF 1        // buffers of reference atoms in shared memory (1: 13 warps per SM fit; 2: double buffered)
#endif
#ifndef RH_QUEUE
#define RH_QUEUE 512      // ring capacity (power of two): the < 64 entries a walk leaves behind plus the 2 * SLOTS * 32 = 384 a
                          // pair of reference atoms can add always fit - no capacity check in front of the push
#endif

struct RdfHalfArgs {
  GridDev g;
  const Rec *R;          // sorted items: .p FP32 (wrapped - L/2) + type code, .u FP64 absolute position + molecule code
  const int32_t *cell_start;
  const int32_t *items;
  const double *thr;     // [nbin + 2]: thr[0..nbin] lower thresholds in s = d^2, thr[nbin + 1] = largest s with sqrt(s) <= last
  int nbin, ntypes, same_molecule, use_smem;
  float first, scale;    // bin estimate (sqrtf(s) - first) * scale
  float bin_delta;       // the FP32 estimate is the exact bin when its fractional part lies in [delta, 1 - delta]
  float r2hi;            // loose FP32 filter: d2 < r2hi may be in range
  unsigned long long *hist;
  int64_t *counters;
  int n;                 // items (checked build: bounds of the sorted array)
};

__global__ void rdf_thresholds_kernel(const double *__restrict__ edges, int nbin, double *__restrict__ thr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nbin + 1) return;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  if (k <= nbin) {       // smallest s >= 0 with sqrt(s) >= edges[k]
    const double e = edges[k];
    double s = e > 0.0 ? e * e : 0.0;
    if (e > 0.0) {
      while (s > 0.0 && sqrt(nextafter(s, -inf)) >= e) s = nextafter(s, -inf);
      while (sqrt(s) < e) s = nextafter(s, inf);
    }
    thr[k] = s;
  } else {               // largest s with sqrt(s) <= edges[nbin]
    const double e = edges[nbin];
    double s = e * e;
    if (e < 0.0) {
      s = -1.0;
    } else {
      while (sqrt(nextafter(s, inf)) <= e) s = nextafter(s, inf);
      while (s > 0.0 && sqrt(s) > e) s = nextafter(s, -inf);
    }
    thr[k] = s;
  }
}

// explicit shared-memory accesses with a 32-bit address held in a register (the compiler otherwise rebuilds the
// shared window base with uniform-datapath instructions at every use inside the hit walk)
__device__ __forceinline__ double4 lds_double4(uint32_t addr) {
  double4 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%4];\nld.shared.v2.f64 {%2, %3}, [%4+16];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_int(uint32_t addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void reds_add(uint32_t addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// SAME_MOL: pairs inside one molecule count (reference sameMolecule=True); ONE_TYPE: a single type-pair key;
// SMEM_HIST: the histogram fits the per-warp shared-memory copy
// Everything one warp owns in shared memory.  The warps of a CTA work independently, each on its own home cells; they
// only share the privatised histogram - 4 KB for 1000 bins, which limits single-warp CTAs to 13 per SM; one histogram
// per two warps gives 16.  (Measured on config 4: 10.1 -> 9.3 ms; a multi-warp CTA pays ~5 % for real WARPSYNCs, which
// the compiler drops when the CTA is one warp, so four warps per CTA gain nothing over two.)
template <int NC>
struct __align__(128) RhWarp {
  Rec land[NC];                       // TMA landing zone; stays valid for the whole pass
  double4 homeU[RH_HOMEBUF][32];      // reference atoms of the home cell, 32 at a time (the next 32 wait in registers
  float4 homeP[RH_HOMEBUF][32];       // until the queue of the current ones is drained)
  int raw[HS_ITEM_INTS];
  unsigned short q[RH_QUEUE];         // hit queue: reference (5 bits) | slot << 5 | lane << 9
  uint64_t bar;
  unsigned qt;                        // ... its tail (running counter, advanced with one atomic per lane)
};

template <int SLOTS, bool SAME_MOL, bool ONE_TYPE, bool SMEM_HIST>
__global__ void __launch_bounds__(32 * RH_CTA_WARPS, RH_WARPS / RH_CTA_WARPS) rdf_sweep_half_kernel(const RdfHalfArgs a) {
  constexpr int NC = SLOTS * 32;
  constexpr int NPAIR = SLOTS / 2;
  extern __shared__ unsigned int s_hist[];            // [npairs * nbin] when use_smem: one copy per CTA
  __shared__ RhWarp<NC> s_warp[RH_CTA_WARPS];
  RhWarp<NC> &W = s_warp[threadIdx.x >> 5];
  Rec(&s_land)[NC] = W.land;
  float4(&s_homeP)[RH_HOMEBUF][32] = W.homeP;
  double4(&s_homeU)[RH_HOMEBUF][32] = W.homeU;
  int(&s_raw)[HS_ITEM_INTS] = W.raw;
  unsigned short(&s_q)[RH_QUEUE] = W.q;
  unsigned &s_qt = W.qt;
  uint64_t &s_bar = W.bar;
  const int lane = threadIdx.x & 31;
  const int nbin = a.nbin;
  const int hsize = a.ntypes * (a.ntypes + 1) / 2 * nbin;
  if (SMEM_HIST)
    for (int k = threadIdx.x; k < hsize; k += 32 * RH_CTA_WARPS) s_hist[k] = 0;
  if (lane == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    s_qt = 0u;
  }
  __syncthreads();
  uint32_t phase = 0;
  const double thr0 = a.thr[0], shi = a.thr[nbin + 1];
  const double Lx = a.g.box[0], Ly = a.g.box[1], Lz = a.g.box[2];
  const double hLx = 0.5 * Lx, hLy = 0.5 * Ly, hLz = 0.5 * Lz;
  const float2 nr2 = make_float2(-a.r2hi, -a.r2hi);
  const uint32_t land_base = smem_addr(s_land);          // candidate (slot c, lane l) = element c * 32 + l
  const uint32_t hist_base = smem_addr(s_hist);

  // fetch pipeline (see p2_sweep_half_kernel): ticket one item ahead, range table one item ahead
  int tkA, tkP = 0;
  {
    int t0 = 0;
    if (lane == 0) {
      t0 = (int)atomicAdd((unsigned int *)&a.counters[2], 1u);
      tkP = (int)atomicAdd((unsigned int *)&a.counters[2], 1u);
    }
    tkA = __shfl_sync(MOUSE_FULL_MASK, t0, 0);
    if (tkA < a.g.ncell) hs_raw_issue(a.items, lane, tkA, s_raw);
  }
  for (;;) {
    HsItem cur;
    bool have = false;
    while (tkA < a.g.ncell) {
      const bool nonempty = hs_finalize(a.g, lane, tkA, s_raw, cur);
      tkA = __shfl_sync(MOUSE_FULL_MASK, tkP, 0);
      if (lane == 0) tkP = (int)atomicAdd((unsigned int *)&a.counters[2], 1u);
      if (tkA < a.g.ncell) hs_raw_issue(a.items, lane, tkA, s_raw);
      if (nonempty) {
        have = true;
        break;
      }
    }
    if (!have) break;
    const int hs = cur.hs, nhome = cur.nhome, T = cur.T;
    const bool bnd = cur.boundary != 0;

    for (int w0 = 0; w0 < T; w0 += NC) {
      // ---- candidates [w0, w1) of the concatenation -> landing zone (TMA) -> FP32 coordinates in registers
      const int w1 = min(T, w0 + NC), nvalid = w1 - w0;
      const int npp = (nvalid + 63) >> 6;
      __syncwarp();          // every lane is done with the previous contents
      fence_proxy_async();
      if (lane == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)nvalid * 48u);
      __syncwarp();
      {
        const int a0 = max(cur.rt, w0), b0 = min(cur.rt + cur.rlen, w1);
        if (b0 > a0) {
          const int src = cur.rsrc + (a0 - cur.rt), cnt = b0 - a0, dst = a0 - w0;
          MOUSE_ASSERT(dst >= 0 && dst + cnt <= NC && src >= 0 && src + cnt <= a.n);
          tma_load_1d(&s_land[dst], a.R + src, (uint32_t)cnt * 48u, &s_bar);
        }
      }
      // the first chunk of reference atoms is fetched while the bulk copies fly
      float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
      double4 hu = make_double4(0.0, 0.0, 0.0, 0.0);
      if (lane < nhome) {
        hp = a.R[hs + lane].p;
        hu = a.R[hs + lane].u;
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
      float2 xr2[NPAIR], yr2[NPAIR], zr2[NPAIR];
      {
#pragma unroll
        for (int pp = 0; pp < NPAIR; ++pp) {
          float q[2][3];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int tl = 32 * (2 * pp + h) + lane;
            const float4 p = s_land[tl].p;
            float px = p.x, py = p.y, pz = p.z;
            if (bnd) hs_nearest_image(px, py, pz, cur.hx, cur.hy, cur.hz, a.g);   // nearest image to the home cell centre
            const bool valid = tl < nvalid;
            q[h][0] = valid ? px : 1.0e18f;
            q[h][1] = valid ? py : 1.0e18f;
            q[h][2] = valid ? pz : 1.0e18f;
          }
          xr2[pp] = make_float2(q[0][0], q[1][0]);
          yr2[pp] = make_float2(q[0][1], q[1][1]);
          zr2[pp] = make_float2(q[0][2], q[1][2]);
        }
      }
      s_homeP[0][lane] = hp;
      s_homeU[0][lane] = hu;
      __syncwarp();

      unsigned qh = s_qt;     // hit queue (ring in shared memory): head (running counter), length - warp-uniform
      int qn = 0;
      // up to two queued hits per lane (entries lane and lane + 32; those >= navail idle): the reference's FP64
      // expression and the exact bin.  The two chains are independent and interleave (the single chain is a long
      // sequence of dependent shared loads, FP64 and conversion instructions)
      auto walk = [&](const int navail, const int buf) {
        bool act[RH_WALK];
        double4 ui[RH_WALK], uj[RH_WALK];
        uint32_t cand[RH_WALK];
        int ilr[RH_WALK];
#pragma unroll
        for (int r = 0; r < RH_WALK; ++r) {
          act[r] = lane + 32 * r < navail;
          const unsigned e = act[r] ? (unsigned)s_q[(qh + (unsigned)(lane + 32 * r)) & (RH_QUEUE - 1)] : 0u;
          const int c = (e >> 5) & 15u, lj = e >> 9;
          MOUSE_ASSERT(c < SLOTS && lj < 32 && (int)(e & 31u) < 32);
          ilr[r] = e & 31u;
          ui[r] = s_homeU[buf][ilr[r]];
          cand[r] = land_base + (uint32_t)(c * 32 + lj) * 48u;
          uj[r] = lds_double4(cand[r] + 16u);
        }
#pragma unroll
        for (int r = 0; r < RH_WALK; ++r) {
          double tx = __dsub_rn(uj[r].x, ui[r].x), ty = __dsub_rn(uj[r].y, ui[r].y), tz = __dsub_rn(uj[r].z, ui[r].z);
          // dr - L * around(dr / L): around() is 0 whenever |dr| <= L / 2 (ordering_functions.py:72-74); one
          // rare branch for coordinates that were not wrapped into the box
          // (measured: a branch-free dr - L * n for EVERY hit costs nine more FP64-pipe operations per hit and makes the
          // sweep 22 % slower; inside the branch, which boundary cells do take, it replaces three divisions)
          if (!(fabs(tx) <= hLx && fabs(ty) <= hLy && fabs(tz) <= hLz)) {
            if (a.g.tri) {      // (inside the brick the triclinic minimum image is the identity, too)
              mouse_min_image_tri(tx, ty, tz, a.g.box, a.g.tilt);
            } else if (fabs(tx) < 2.5 * hLx && fabs(ty) < 2.5 * hLy && fabs(tz) < 2.5 * hLz) {
              // around(dr / L) is +-1 for L / 2 < |dr| < 5 L / 4 (dr > L/2 means dr >= L/2 (1 + 2^-52): the rounded
              // quotient is >= 0.5 + 2^-53 > 0.5; exactly L / 2 rounds half-to-even to 0) and 0 below: dr - L * n with
              // an exact L * n and one rounded subtraction, no division (pairs across the periodic boundary)
              tx = __dsub_rn(tx, tx > hLx ? Lx : (tx < -hLx ? -Lx : 0.0));
              ty = __dsub_rn(ty, ty > hLy ? Ly : (ty < -hLy ? -Ly : 0.0));
              tz = __dsub_rn(tz, tz > hLz ? Lz : (tz < -hLz ? -Lz : 0.0));
            } else {
              tx = mouse_trim(tx, Lx);
              ty = mouse_trim(ty, Ly);
              tz = mouse_trim(tz, Lz);
            }
          }
          const double s = __dadd_rn(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)), __dmul_rn(tz, tz));
          // np.ma.masked_equal(d, 0.) drops coincident atoms (:84); values outside [first, last] are dropped (:86)
          bool keep = act[r] && s != 0.0 && s >= thr0 && s <= shi;
          if (!SAME_MOL) keep = keep && uj[r].w != ui[r].w;
          // bin estimate from an FP32 square root (error << one bin), made exact with the two thresholds around it
          const float sf = fmaxf((float)s, 1.0e-30f);
          float rs;
          asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(sf));      // one MUFU.RSQ; 2 ulp is plenty here
          const float v = (sf * rs - a.first) * a.scale;
          int k = (int)v;
          const float fr = v - (float)k;
          // the estimate is off by far less than bin_delta bins: away from the edges it IS the bin, and the two
          // thresholds (global loads) are only fetched for the few hits next to an edge
          const bool sure = (fr > a.bin_delta) & (fr < 1.0f - a.bin_delta) & (k >= 0) & (k < nbin);
          k = min(max(k, 0), nbin - 1);
          if (keep && !sure) {
            const double tlo = __ldg(a.thr + k), thi = __ldg(a.thr + k + 1);
            k += (s >= thi && k < nbin - 1) ? 1 : 0;
            k -= (s < tlo && k > 0) ? 1 : 0;
          }
          int slot = 0;
          if (!ONE_TYPE) {
            const int ti = __float_as_int(s_homeP[buf][ilr[r]].w);
            const int tj = lds_int(cand[r] + 12u);
            const int ta = min(ti, tj), tb = max(ti, tj);
            slot = (ta * a.ntypes - ta * (ta - 1) / 2 + (tb - ta)) * nbin;
          }
          if (keep) {
            MOUSE_ASSERT(k >= 0 && k < nbin && slot >= 0 && slot + k < hsize);
            if (SMEM_HIST) reds_add(hist_base + 4u * (uint32_t)(slot + k), 2u);
            else atomicAdd(&a.hist[(size_t)(slot + k)], 2ULL);
          }
        }
      };
      for (int chunk0 = 0, buf = 0; chunk0 < nhome; chunk0 += 32, buf ^= (RH_HOMEBUF - 1)) {
        const int nchunk = min(32, nhome - chunk0);
        // next chunk of reference atoms: loaded now, stored after this chunk has been processed
        if (chunk0 + 32 + lane < nhome) {
          hp = a.R[hs + chunk0 + 32 + lane].p;
          hu = a.R[hs + chunk0 + 32 + lane].u;
        }
        for (int il = 0; il < nchunk; il += 2) {
          // two reference atoms per iteration (two independent filter chains, one queue update); the second one
          // of an odd tail is switched off
          const float4 pa = s_homeP[buf][il], pb = s_homeP[buf][(il + 1) & 31];
          const bool two = il + 1 < nchunk;
          // ---- FP32 filter: slot c -> bit c of the lane's mask
          const float2 nax = make_float2(-pa.x, -pa.x), nay = make_float2(-pa.y, -pa.y), naz = make_float2(-pa.z, -pa.z);
          const float2 nbx = make_float2(-pb.x, -pb.x), nby = make_float2(-pb.y, -pb.y), nbz = make_float2(-pb.z, -pb.z);
          unsigned maska = 0, maskb = 0;
#pragma unroll
          for (int pp = NPAIR - 1; pp >= 0; --pp) {
            if (pp < npp) {
              const float2 dxa = __fadd2_rn(xr2[pp], nax), dya = __fadd2_rn(yr2[pp], nay), dza = __fadd2_rn(zr2[pp], naz);
              const float2 dxb = __fadd2_rn(xr2[pp], nbx), dyb = __fadd2_rn(yr2[pp], nby), dzb = __fadd2_rn(zr2[pp], nbz);
              const float2 ta = __fadd2_rn(__ffma2_rn(dza, dza, __ffma2_rn(dya, dya, __fmul2_rn(dxa, dxa))), nr2);   // < 0: maybe in range
              const float2 tb = __fadd2_rn(__ffma2_rn(dzb, dzb, __ffma2_rn(dyb, dyb, __fmul2_rn(dxb, dxb))), nr2);
              maska = __funnelshift_l(__float_as_uint(ta.y), maska, 1);
              maska = __funnelshift_l(__float_as_uint(ta.x), maska, 1);
              maskb = __funnelshift_l(__float_as_uint(tb.y), maskb, 1);
              maskb = __funnelshift_l(__float_as_uint(tb.x), maskb, 1);
            }
          }
          if (w0 < nhome) {       // home-cell candidates lead the concatenation: count home-home pairs once (t > i)
            const int lim = chunk0 + il - w0 - lane;        // slot c allowed iff 32 c > lim
            const int cmina = min(max((lim + 32) >> 5, 0), 32), cminb = min(max((lim + 33) >> 5, 0), 32);
            maska &= cmina >= 32 ? 0u : ~((1u << cmina) - 1u);
            maskb &= cminb >= 32 ? 0u : ~((1u << cminb) - 1u);
          }
          if (!two) maskb = 0u;
          // ---- the hits go to the warp's queue (reference, slot, lane), so that the exact FP64 part below runs with
          // two hits per lane whatever their distribution over the lanes (a lane holds 0..SLOTS of them per reference)
          auto push = [&](unsigned ma, unsigned mb) {
            const int n2 = __popc(ma) + __popc(mb);
            if (n2) {
              unsigned pos = atomicAdd(&s_qt, (unsigned)n2);
              unsigned ebase = (unsigned)il | ((unsigned)lane << 9);
              while (ma) {
                const int c = 31 - __clz((int)ma);
                ma &= ~(1u << c);
                s_q[pos & (RH_QUEUE - 1)] = (unsigned short)(ebase | ((unsigned)c << 5));
                ++pos;
              }
              ++ebase;
              while (mb) {
                const int c = 31 - __clz((int)mb);
                mb &= ~(1u << c);
                s_q[pos & (RH_QUEUE - 1)] = (unsigned short)(ebase | ((unsigned)c << 5));
                ++pos;
              }
            }
          };
          auto drain = [&]() {
            __syncwarp();
            while (qn > 0) {
              const int m = min(qn, 32 * RH_WALK);
              walk(m, buf);
              qh += m;
              qn -= m;
            }
            __syncwarp();
          };
          if (RH_QUEUE >= 64 + 2 * NC) {        // (compile time) the ring cannot overflow: the new length is read back
            push(maska, maskb);
            __syncwarp();
            qn = (int)(*(volatile unsigned *)&s_qt - qh);
          } else {
            int tot = __reduce_add_sync(MOUSE_FULL_MASK, __popc(maska) + __popc(maskb));
            if (qn + tot > RH_QUEUE) {      // cannot happen below ~3 hits per lane and reference
              drain();
              if (tot > RH_QUEUE) {         // more than 4 per lane and reference: one reference at a time
                const int ta = __reduce_add_sync(MOUSE_FULL_MASK, __popc(maska));
                push(maska, 0u);
                qn = ta;
                drain();
                maska = 0u;
                tot -= ta;
              }
            }
            push(maska, maskb);
            qn += tot;
            __syncwarp();
          }
          // ---- full rounds of 64 queued hits: exact FP64 distance, exact bin, two hits per lane
          while (qn >= 32 * RH_WALK) {
            walk(32 * RH_WALK, buf);
            qh += 32 * RH_WALK;
            qn -= 32 * RH_WALK;
          }
          __syncwarp();
        }
        if (qn > 0) {       // the reference chunk (and, after the last chunk, the landing zone) is about to change
          walk(qn, buf);
          qh += qn;
          qn = 0;
        }
        __syncwarp();
        s_homeP[buf ^ (RH_HOMEBUF - 1)][lane] = hp;
        s_homeU[buf ^ (RH_HOMEBUF - 1)][lane] = hu;
        __syncwarp();
      }
    }
  }
  if (SMEM_HIST) {
    __syncthreads();        // every warp of the CTA has run out of home cells
    for (int k = threadIdx.x; k < hsize; k += 32 * RH_CTA_WARPS)
      if (s_hist[k]) atomicAdd(&a.hist[k], (unsigned long long)s_hist[k]);
  }
}

template <bool SAME_MOL, bool ONE_TYPE, bool SMEM_HIST>
static void launch_rdf_half(const RdfHalfArgs &h, int ncell, size_t smem, cudaStream_t st) {
  // (no cached state: the library holds nothing per process; the query costs microseconds next to a 10 ms sweep)
  int blocks_per_sm = 0;
  // static + dynamic shared memory passes 48 KB with the histogram: opt in (per device and function, cheap)
  cudaFuncSetAttribute(rdf_sweep_half_kernel<RH_SLOTS, SAME_MOL, ONE_TYPE, SMEM_HIST>,
                       cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, rdf_sweep_half_kernel<RH_SLOTS, SAME_MOL, ONE_TYPE, SMEM_HIST>,
                                                32 * RH_CTA_WARPS, smem);
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  const int sms = mouse_num_sms();
  int blocks = sms * blocks_per_sm;
  if (blocks > (ncell + RH_CTA_WARPS - 1) / RH_CTA_WARPS) blocks = (ncell + RH_CTA_WARPS - 1) / RH_CTA_WARPS;
  rdf_sweep_half_kernel<RH_SLOTS, SAME_MOL, ONE_TYPE, SMEM_HIST><<<blocks, 32 * RH_CTA_WARPS, smem, st>>>(h);
}

int mouse_launch_rdf_sweep(const mouse_grid_t *grid, uint32_t flags, int64_t n, int32_t ntypes,
                           const double *edges_dev, int32_t nbin, const Workspace &w, int64_t *hist,
                           cudaStream_t st) {
  const size_t hsize = (size_t)ntypes * (ntypes + 1) / 2 * (size_t)nbin;
  MOUSE_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int64_t) * hsize, st));
  if (n == 0) return MOUSE_OK;
  const bool fast = grid->fast && !(flags & MOUSE_FORCE_EXACT);
  if (fast) {
    RdfHalfArgs h;
    h.g = mouse_grid_dev(grid);
    h.R = w.rec_sorted;
    h.cell_start = w.cell_start;
    h.items = w.items;
    h.thr = w.thr;
    h.nbin = nbin;
    h.ntypes = ntypes;
    h.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
    h.hist = (unsigned long long *)hist;
    h.counters = w.counters;
    h.n = (int)n;
    // bin estimate and loose filter need the range on the host: grid->rcut is the last edge (engine.rdf_frame),
    // the first edge rides in through rdf_first
    h.first = (float)w.rdf_first;
    h.scale = (float)((double)nbin / (grid->rcut - w.rdf_first));
    h.r2hi = nextafterf((float)(grid->rc2 * (1.0 + (double)grid->guard)), 3.0e38f);
    // error of the FP32 bin estimate in bins: sqrt through (float)s * rsqrt.approx (relative < 4e-7 of a value <= last),
    // the subtraction and the scaling (a few ulps of a value <= nbin); delta = 3 x that (already generous) bound, at
    // least 1e-3
    {
      const double width = (grid->rcut - w.rdf_first) / (double)nbin;
      double d = 3.0 * (4.0e-7 * grid->rcut / width + 3.0e-7 * (double)nbin);
      if (d < 1.0e-3) d = 1.0e-3;
      h.bin_delta = (float)d;      // (>= 0.5: every hit takes the exact check)
    }
    size_t smem = hsize * sizeof(unsigned int);
    h.use_smem = smem <= 16 * 1024;
    if (!h.use_smem) smem = 0;
    rdf_thresholds_kernel<<<(nbin + 2 + 127) / 128, 128, 0, st>>>(edges_dev, nbin, w.thr);
    MOUSE_LAUNCH_CHECK("rdf_thresholds_kernel");
    const int key = (h.same_molecule ? 4 : 0) | (ntypes == 1 ? 2 : 0) | (h.use_smem ? 1 : 0);
    switch (key) {
      case 0: launch_rdf_half<false, false, false>(h, grid->ncell, smem, st); break;
      case 1: launch_rdf_half<false, false, true>(h, grid->ncell, smem, st); break;
      case 2: launch_rdf_half<false, true, false>(h, grid->ncell, smem, st); break;
      case 3: launch_rdf_half<false, true, true>(h, grid->ncell, smem, st); break;
      case 4: launch_rdf_half<true, false, false>(h, grid->ncell, smem, st); break;
      case 5: launch_rdf_half<true, false, true>(h, grid->ncell, smem, st); break;
      case 6: launch_rdf_half<true, true, false>(h, grid->ncell, smem, st); break;
      default: launch_rdf_half<true, true, true>(h, grid->ncell, smem, st); break;
    }
    MOUSE_LAUNCH_CHECK("rdf_sweep_half_kernel");
    return MOUSE_OK;
  }
  RdfArgs a;
  a.g = mouse_grid_dev(grid);
  a.R = w.rec_sorted;
  a.cell_start = w.cell_start;
  a.cell_sorted = w.cell_sorted;
  a.edges = edges_dev;
  a.nbin = nbin;
  a.ntypes = ntypes;
  a.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
  a.hist = (unsigned long long *)hist;
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
