// K3: the pair sweep.  Replaces calculateAveCosSqForReference (ordering_functions.py:91-119)
// for all reference bonds of a frame at once.
//
//  * p2_sweep_half_kernel - the fast path.  Half shell (Newton's third law): every unordered pair of
//    bonds is decided ONCE - home cell against itself (j > i) and against its 13 "forward" neighbour
//    cells - and its cos^2 is added to both partners.  A work item is one ITEM ROW (half_shell.cuh): a home cell, or
//    one window of 32 references of a crowded home cell; one warp per item, persistent warps with a ticket counter,
//    rows ordered heaviest cells first.  The candidates of the 14 cells are up to 10 contiguous ranges of the
//    cell-sorted arrays: they are fetched with 1-D TMA bulk copies (cp.async.bulk + mbarrier) into a shared-memory
//    landing zone (a chunk of 256 records), culled (hs_cull: candidates farther than the cutoff from the bounding box
//    of the home cell's bonds are dropped, 220 -> 148 per reference on the 1M-bond melt), moved into REGISTERS (lane
//    t % 32, slot t / 32: FP32 wrapped coordinates as packed pairs, FP64 unit vectors), and the landing zone is
//    immediately re-armed with the NEXT chunk / item, whose copies then overlap the arithmetic.  The references are
//    taken in groups of eight, two per loop iteration (warp-uniform loop, broadcast LDS, one loop body per number of
//    slots in use): packed f32x2 distance test against the guard band, dense FP64 cos^2 = (u_i . u_j)^2 for every
//    slot, branch-free accumulation into the reference's and the candidate's registers (the hit mask - the sign of
//    d2 - rc2_lo - clears the weight and is subtracted from the counts).  Candidates inside the FP32 guard band around
//    rc^2, and sure hits with |u_i . u_j| < 1e-9 (the reference masks cos^2 == 0, :113), are NOT decided in FP32:
//    they are appended to a pair list and decided with the reference's exact un-fused FP64 expressions by
//    p2_fixup_kernel.  Sums are merged across items as 2^-47 fixed point with integer atomics: exact, order
//    independent -> run-to-run bit-identical.
//  * p2_sweep_exact_kernel - one thread per reference, all FP64, literally the reference's
//    expression per candidate.  Valid for ANY grid (fewer than 3 cells on an axis, rcut < 0,
//    rmax == 0 -> half diagonal, more than 512 items per cell); it also extracts the pair lists, and its body is the
//    automatic fallback (inside the fixup launch) when the pair list of the fast path overflows.
#include <type_traits>

#include "half_shell.cuh"

struct SweepArgs {
  GridDev g;
  const Rec *R;            // sorted items: .p FP32 (wrapped midpoint - L/2) + molecule code bits, .u FP64 unit bond
                           // vector with .w = bit pattern of the sorted slot index
  const int32_t *idx;
  const int32_t *cell_start;
  const int32_t *items;    // per-cell item table of the half-shell sweep (half_shell.cuh)
  const int32_t *cell_sorted;
  const uint8_t *flag;
  const double *vec;       // SoA [3][n], original order
  const double *ctr;
  int n;
  double *out_sum;
  int32_t *out_cnt;
  int64_t *counters;       // 0 zero-length bonds, 1 exact re-decisions (diagnostic), 2 home-cell ticket,
                           // 3 length of the pair list, 4 pair list overflowed
  const int64_t *row_ptr;  // lists only
  int32_t *nbr_idx;
  int64_t list_cap;        // lists only: entries of nbr_idx the caller allocated
  int2 *fix_list;          // pairs of sorted slots left to p2_fixup_kernel: x = i | kind << 31, y = j
  long long fix_cap;
  long long *acc_fix;      // [n] per sorted slot: sum of cos^2 as 2^-47 fixed point
  int32_t *acc_cnt;        // [n] per sorted slot: counted neighbours
  int same_molecule;
  int no_cutoff;
};

#define C_NZERO 0
#define C_NFIX_DONE 1
#define C_TICKET 2
#define C_NFIX 3
#define C_OVERFLOW 4

// ------------------------------------------------------------------ rare exact path
// The reference's decision for ONE pair of sorted slots (si = reference, sj = candidate): returns
// cos^2 exactly as ordering_functions.py:95-113 evaluates it, or 0.0 when the pair is not counted
// (out of range, or cos^2 == 0 which the reference masks).
__device__ __forceinline__ double exact_pair(const int32_t *__restrict__ idx, const double *__restrict__ ctr,
                                             const double *__restrict__ vec, size_t n, const GridDev &g, int si,
                                             int sj) {
  const int bi = idx[si], bj = idx[sj];
  const double rc2 = g.rc2;
  const double rsq = mouse_rsq_exact_g(ctr[bi], ctr[n + bi], ctr[2 * n + bi], ctr[bj], ctr[n + bj], ctr[2 * n + bj],
                                       g.box, g.tilt, g.tri);
  if (!(rsq <= rc2)) return 0.0;
  const double ax = vec[bi], ay = vec[n + bi], az = vec[2 * n + bi];
  const double bx = vec[bj], by = vec[n + bj], bz = vec[2 * n + bj];
  const double dot = mouse_dot_exact(ax, ay, az, bx, by, bz);
  const double nr = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
  const double nj = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
  return __ddiv_rn(__ddiv_rn(__dmul_rn(dot, dot), nr), nj);
}

// ------------------------------------------------------------------ half-shell fast sweep
#ifndef HS_MAXHOME
#define HS_MAXHOME MOUSE_WINDOW   // references per work item: a home cell with more bonds has one item row per window
#endif
#ifndef HS_GRP
#define HS_GRP 8           // references reduced together (transposed reduction through shared memory)
#endif
#ifndef HS_LZ
#define HS_LZ 256          // landing zone of the TMA copies, in records: one chunk of the candidate concatenation
#endif
#define HS_MAXT 16384      // more candidates than this per home cell: exact path (keeps the 2^-47 fixed point in range)
#ifndef HS_WARPS
#define HS_WARPS 10        // resident warps (= CTAs) per SM the register allocation is sized for
#endif
#define HS_FIX_SCALE 140737488355328.0     // 2^47
#define HS_FIX_INV 7.105427357601002e-15   // 2^-47
#define HS_NEARZ 1e-9                      // |u_i . u_j| below this is decided exactly (:113)
#define HS_NEARZ_HI 0x3e112e0b             // high word of 1e-9
#define HS_PAD 1.0e18f                     // coordinate of an empty slot

// Pairs the sweep does not decide in FP32 are appended to a global pair list and decided by p2_fixup_kernel
// (no function call and no extra state inside the sweep: its register budget is spent on the candidates).
__device__ __forceinline__ void fix_append(int64_t *counters, int2 *fix_list, long long fix_cap, int si, int sj,
                                           int kind) {
  const unsigned long long pos = atomicAdd((unsigned long long *)&counters[C_NFIX], 1ULL);
  if (pos < (unsigned long long)fix_cap) fix_list[pos] = make_int2(si | (kind << 31), sj);
  else counters[C_OVERFLOW] = 1;   // the exact sweep recomputes the whole frame
}

// every pair of a home cell with more candidates than the fixed-point accumulators are sized for goes to the pair list
// (never reached through mouse_p2_uses_fast: such frames take the all-FP64 sweep; kept as a safety net)
__device__ __noinline__ void hs_oversize(const Rec *__restrict__ R, int64_t *counters, int2 *fix_list,
                                         long long fix_cap, int same_molecule, int lane, int hs, int nhome, int rsrc,
                                         int rlen, int rt0) {
  for (int r = 0; r < 10; ++r) {
    const int src = __shfl_sync(MOUSE_FULL_MASK, rsrc, r), len = __shfl_sync(MOUSE_FULL_MASK, rlen, r);
    const int rt = __shfl_sync(MOUSE_FULL_MASK, rt0, r);
    // the frame is being redone exactly anyway once the list has overflowed (warp-uniform decision: the lanes must
    // stay together for the shuffles above and in the caller)
    if (__any_sync(MOUSE_FULL_MASK, ((volatile int64_t *)counters)[C_OVERFLOW] != 0)) return;
    for (int k = lane; k < len; k += 32) {
      const int sj = src + k, tj = rt + k;
      const int molj = __float_as_int(R[sj].p.w);
      for (int il = 0; il < nhome; ++il) {
        if (tj <= il) continue;                                   // home-home pairs once (j > i)
        if (!same_molecule && __float_as_int(R[hs + il].p.w) == molj) continue;
        fix_append(counters, fix_list, fix_cap, hs + il, sj, 0);
      }
    }
  }
}

// @(cnt > 0): acc_fix[g] += round(acc * 2^47); acc_cnt[g] += cnt   (no return value: RED)
__device__ __forceinline__ void hs_flush(long long *acc_fix, int32_t *acc_cnt, int g, double acc, int cnt) {
  const long long v = __double2ll_rn(acc * HS_FIX_SCALE);
  asm volatile("{\n.reg .pred p;\nsetp.gt.s32 p, %3, 0;\n@p red.global.add.u64 [%0], %2;\n@p red.global.add.s32 [%1], %3;\n}" ::"l"(
                   acc_fix + g),
               "l"(acc_cnt + g), "l"(v), "r"(cnt)
               : "memory");
}

// explicit shared-memory accesses with a 32-bit address held in a register
__device__ __forceinline__ float4 lds_float4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_float4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ double2 lds_double2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return (uint32_t)v;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, int v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

#ifndef HS_REF_UNROLL
#define HS_REF_UNROLL 1       // trips of the reference-pair loop unrolled together
#endif
constexpr int hs_ref_unroll = HS_REF_UNROLL;

struct HsConst {
  float nlo;            // -rc2_lo: t = d2 - rc2_lo is negative for a sure hit
  unsigned band_bits;   // bit pattern of the guard band width: 0 <= t <= width (compared as unsigned) is undecided
  float nearz_thr;
};

// One reference of the current window as the loop reads it (48 bytes, three 128-bit broadcast loads): the negated FP32
// coordinates duplicated for the packed f32x2 subtraction and the FP64 unit vector (the molecule codes live in s_mol).
struct __align__(16) HsRef {
  float4 a;    // -x, -x, -y, -y
  float2 b;    // -z, -z
  double uz;
  double2 u;   // ux, uy
};
#define HS_REF_BYTES 48u

__device__ __forceinline__ unsigned hs_umin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u32(a, b, c); }

// Two packed floats in ONE 64-bit register (an aligned register pair by construction: the packed f32x2 instructions
// then take their operands as they are, the compiler never has to re-pair two scalar registers inside the loop).
typedef unsigned long long pk2;
__device__ __forceinline__ pk2 pk2_make(float lo, float hi) {
  pk2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void pk2_bits(pk2 v, int &lo, int &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ void pk2_floats(pk2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pk2 pk2_add(pk2 a, pk2 b) {
  pk2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk2 pk2_mul(pk2 a, pk2 b) {
  pk2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ pk2 pk2_fma(pk2 a, pk2 b, pk2 c) {
  pk2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 16 bytes of shared memory as two packed pairs / as two doubles
__device__ __forceinline__ void lds_pk2x2(uint32_t addr, pk2 &a, pk2 &b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
// @(m < 0): accj += d * d; acci += d * d   (the hit mask m predicates the FP64 accumulation: no select, no copy of d)
__device__ __forceinline__ void hs_acc_if(double &accj, double &acci, double d, int m) {
  asm("{\n.reg .pred p;\nsetp.lt.s32 p, %3, 0;\n@p fma.rn.f64 %0, %2, %2, %0;\n@p fma.rn.f64 %1, %2, %2, %1;\n}"
      : "+d"(accj), "+d"(acci)
      : "d"(d), "r"(m));
}

// One group of references (nref <= HS_GRP, two per iteration) against NS register-resident slots = NPP slot pairs
// (specialised on NS: no guards inside the loop, the slot pairs interleave freely; with an odd NS the FP64 work of the
// last pair's empty second slot is left out).  Returns the rare mask: two bits per pair of references.  Only this loop
// is instantiated per slot count - the code around it exists once (the kernel's instruction footprint matters: with
// every slot count inlining its own copy of the rare path and of the reduction the kernel ran 20 % slower).
template <int SLOTS, int NS, bool EXCL_MOL>
__device__ __forceinline__ unsigned hs_ref_pairs(const HsConst k, const int lane, const int base, const int nref,
                                                 const int (&jl)[HS_MAXHOME / 32], const pk2 (&xr2)[SLOTS / 2],
                                                 const pk2 (&yr2)[SLOTS / 2], const pk2 (&zr2)[SLOTS / 2],
                                                 const double (&ux)[SLOTS], const double (&uy)[SLOTS],
                                                 const double (&uz)[SLOTS], double (&accj)[SLOTS], int (&cntj)[SLOTS],
                                                 const int (&molr)[SLOTS], const uint32_t s_ref, const int *s_mol,
                                                 double (*s_acc)[33], uint8_t (*s_cnt)[36]) {
  constexpr int NPP = (NS + 1) / 2;
  constexpr int HSL = HS_MAXHOME / 32;   // slots of pass 0 that may hold home-cell bonds
  const pk2 nlo = pk2_make(k.nlo, k.nlo);
  unsigned rare = 0u;
  // two references per iteration: two independent dependency chains per warp.  The second one of an odd
  // tail is the sentinel record behind the window (far away: no hits, never in the band)
  uint32_t rp = s_ref + HS_REF_BYTES * (uint32_t)base;
  uint32_t ap = smem_addr(&s_acc[0][lane]), cp = smem_addr(&s_cnt[0][lane]);      // (counts: one byte per lane, <= SLOTS)
#pragma unroll hs_ref_unroll
  for (int r = 0; r < nref; r += 2, rp += 2u * HS_REF_BYTES, ap += 2u * 33u * 8u, cp += 2u * 36u) {
    const int il = base + r;
    pk2 nax, nay, naz, nbx, nby, nbz, uaz2, ubz2;
    lds_pk2x2(rp, nax, nay);
    lds_pk2x2(rp + 16u, naz, uaz2);
    lds_pk2x2(rp + HS_REF_BYTES, nbx, nby);
    lds_pk2x2(rp + HS_REF_BYTES + 16u, nbz, ubz2);
    const double2 ua = lds_double2(rp + 32u), ub = lds_double2(rp + HS_REF_BYTES + 32u);
    const double uax = ua.x, uay = ua.y, uaz = __longlong_as_double((long long)uaz2);
    const double ubx = ub.x, uby = ub.y, ubz = __longlong_as_double((long long)ubz2);
    int mola = 0, molb = 0;
    if (EXCL_MOL) {
      mola = s_mol[il];
      molb = s_mol[il + 1];
    }
    double acca0 = 0.0, acca1 = 0.0, accb0 = 0.0, accb1 = 0.0;
    int cnta = 0, cntb = 0;
    unsigned banda = 0xffffffffu, bandb = 0xffffffffu;
    float nza = 3.0e38f, nzb = 3.0e38f;
#pragma unroll
    for (int pp = 0; pp < NPP; ++pp) {
      const pk2 dxa = pk2_add(xr2[pp], nax), dya = pk2_add(yr2[pp], nay), dza = pk2_add(zr2[pp], naz);
      const pk2 dxb = pk2_add(xr2[pp], nbx), dyb = pk2_add(yr2[pp], nby), dzb = pk2_add(zr2[pp], nbz);
      const pk2 ta = pk2_add(pk2_fma(dza, dza, pk2_fma(dya, dya, pk2_mul(dxa, dxa))), nlo);   // d2 - rc2_lo
      const pk2 tb = pk2_add(pk2_fma(dzb, dzb, pk2_fma(dyb, dyb, pk2_mul(dxb, dxb))), nlo);
      int ta0, ta1, tb0, tb1;
      pk2_bits(ta, ta0, ta1);
      pk2_bits(tb, tb0, tb1);
      // smallest non-negative t (negative values are huge as unsigned): in the guard band if <= its width
      banda = hs_umin3(banda, (unsigned)ta0, (unsigned)ta1);
      bandb = hs_umin3(bandb, (unsigned)tb0, (unsigned)tb1);
      const bool two = 2 * pp + 1 < NS;     // (compile time) the pair's second slot is in use
      const double da0 = fma(uax, ux[2 * pp], fma(uay, uy[2 * pp], uaz * uz[2 * pp]));
      const double db0 = fma(ubx, ux[2 * pp], fma(uby, uy[2 * pp], ubz * uz[2 * pp]));
      // Branch-free: the hit mask is the sign of t spread over the word (combined with the home-cell j > i rule and
      // the molecule test); the weight w is d with its high word cleared on a miss (|w| < 1e-300, w * w vanishes in
      // the sums), so the FP64 FMAs are unconditional, and the counts subtract the masks
      int ma0 = ta0 >> 31, mb0 = tb0 >> 31;
      if (2 * pp < HSL) {
        ma0 &= (il - jl[2 * pp]) >> 31;
        mb0 &= (il + 1 - jl[2 * pp]) >> 31;
      }
      if (EXCL_MOL) {
        ma0 &= molr[2 * pp] != mola ? -1 : 0;
        mb0 &= molr[2 * pp] != molb ? -1 : 0;
      }
      const double wa0 = __hiloint2double(__double2hiint(da0) & ma0, __double2loint(da0));
      const double wb0 = __hiloint2double(__double2hiint(db0) & mb0, __double2loint(db0));
      accj[2 * pp] = fma(wb0, wb0, fma(wa0, wa0, accj[2 * pp]));
      acca0 = fma(wa0, wa0, acca0);
      accb0 = fma(wb0, wb0, accb0);
      cntj[2 * pp] = cntj[2 * pp] - ma0 - mb0;
      if (two) {
        const double da1 = fma(uax, ux[2 * pp + 1], fma(uay, uy[2 * pp + 1], uaz * uz[2 * pp + 1]));
        const double db1 = fma(ubx, ux[2 * pp + 1], fma(uby, uy[2 * pp + 1], ubz * uz[2 * pp + 1]));
        // smallest |d| over ALL slots (as the high word, compared as a float): a false alarm only costs time
        nza = fminf(nza, fminf(fabsf(__int_as_float(__double2hiint(da0))), fabsf(__int_as_float(__double2hiint(da1)))));
        nzb = fminf(nzb, fminf(fabsf(__int_as_float(__double2hiint(db0))), fabsf(__int_as_float(__double2hiint(db1)))));
        int ma1 = ta1 >> 31, mb1 = tb1 >> 31;
        if (2 * pp + 1 < HSL) {
          ma1 &= (il - jl[2 * pp + 1]) >> 31;
          mb1 &= (il + 1 - jl[2 * pp + 1]) >> 31;
        }
        if (EXCL_MOL) {
          ma1 &= molr[2 * pp + 1] != mola ? -1 : 0;
          mb1 &= molr[2 * pp + 1] != molb ? -1 : 0;
        }
        const double wa1 = __hiloint2double(__double2hiint(da1) & ma1, __double2loint(da1));
        const double wb1 = __hiloint2double(__double2hiint(db1) & mb1, __double2loint(db1));
        accj[2 * pp + 1] = fma(wb1, wb1, fma(wa1, wa1, accj[2 * pp + 1]));
        acca1 = fma(wa1, wa1, acca1);
        accb1 = fma(wb1, wb1, accb1);
        cntj[2 * pp + 1] = cntj[2 * pp + 1] - ma1 - mb1;
        cnta = cnta - ma0 - ma1;
        cntb = cntb - mb0 - mb1;
      } else {
        nza = fminf(nza, fabsf(__int_as_float(__double2hiint(da0))));
        nzb = fminf(nzb, fabsf(__int_as_float(__double2hiint(db0))));
        cnta -= ma0;
        cntb -= mb0;
      }
    }
    // either reference of the pair has a candidate in the band or a near-zero dot: both are re-tested below
    if ((min(banda, bandb) <= k.band_bits) | (fminf(nza, nzb) <= k.nearz_thr)) rare |= 3u << r;
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(ap), "d"(acca0 + acca1) : "memory");
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(ap + 33u * 8u), "d"(accb0 + accb1) : "memory");
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(cp), "r"(cnta) : "memory");
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(cp + 36u), "r"(cntb) : "memory");
  }
  return rare;
}

// The references of one home cell (window) against the register-resident candidates: groups of HS_GRP references
// through hs_ref_pairs, then - once per group - the rare path and the reduction of the reference-side partials.
template <int SLOTS, bool EXCL_MOL>
__device__ __forceinline__ void hs_ref_loop(const SweepArgs &a, const HsConst k, const int lane, const bool pass0,
                                            const int ns, const int hs, const int nhome, const pk2 (&xr2)[SLOTS / 2],
                                            const pk2 (&yr2)[SLOTS / 2], const pk2 (&zr2)[SLOTS / 2],
                                            const double (&ux)[SLOTS], const double (&uy)[SLOTS],
                                            const double (&uz)[SLOTS], double (&accj)[SLOTS], int (&cntj)[SLOTS],
                                            const int (&molr)[SLOTS], const uint32_t s_ref, const int *s_mol, double (*s_acc)[33],
                                            uint8_t (*s_cnt)[36], const int (*s_gid)[32]) {
  static_assert(HS_GRP == 8 || HS_GRP == 4, "reference groups of 4 or 8");
  static_assert(HS_MAXHOME % 2 == 0, "references are processed two at a time");
  constexpr int HSL = HS_MAXHOME / 32;   // slots of pass 0 that may hold home-cell bonds
  // home-home pairs are decided once: candidate t = 32 c + lane of pass 0 counts for reference il only if t > il.
  // (il - t) >> 31 is the all-ones mask of that rule; jl[c] = t (pass 0) or a large number (every il passes)
  int jl[HSL];
#pragma unroll
  for (int c = 0; c < HSL; ++c) jl[c] = pass0 ? 32 * c + lane : 0x3fffffff;
  // References are processed in groups of HS_GRP.  The inner loop is one basic block: the rare-candidate
  // decision is only recorded (two bits per pair of references) and the reference-side partial of reference il is
  // stored while reference il + 1 is computed, so no iteration ends by draining its own FP64 chain.
  for (int base = 0; base < nhome; base += HS_GRP) {
    const int nref = min(HS_GRP, nhome - base);
    unsigned rare;
#define HS_PAIRS(N)                                                                                                   \
  rare = hs_ref_pairs<SLOTS, (N <= SLOTS ? N : SLOTS), EXCL_MOL>(k, lane, base, nref, jl, xr2, yr2, zr2, ux, uy, uz, accj, cntj, \
                                                                  molr, s_ref, s_mol, s_acc, s_cnt)
    // (a tree of warp-uniform compares: the jump table of a switch costs an indirect branch per group)
    if (SLOTS > 6 && ns > 6) {
      if (ns > 7) {
        if constexpr (SLOTS >= 8) HS_PAIRS(8);
      } else {
        if constexpr (SLOTS >= 7) HS_PAIRS(7);
      }
    } else if (ns > 3) {
      if (ns > 5) {
        if constexpr (SLOTS >= 6) HS_PAIRS(6);
      } else if (ns > 4) {
        if constexpr (SLOTS >= 5) HS_PAIRS(5);
      } else {
        if constexpr (SLOTS >= 4) HS_PAIRS(4);
      }
    } else if (ns > 1) {
      if (ns > 2) {
        if constexpr (SLOTS >= 3) HS_PAIRS(3);
      } else {
        HS_PAIRS(2);
      }
    } else {
      HS_PAIRS(1);
    }
#undef HS_PAIRS
    rare &= (1u << nref) - 1u;     // the sentinel of an odd tail is not a reference
    // ---- rare: candidates inside the guard band / near-zero dots are handed to p2_fixup_kernel
    while (rare) {
      const int r = __ffs(rare) - 1;
      rare &= rare - 1;
      const int il = base + r;
      const uint32_t rq = s_ref + HS_REF_BYTES * (uint32_t)il;
      const float4 ria = lds_float4(rq);
      pk2 riz2, uiz2;
      lds_pk2x2(rq + 16u, riz2, uiz2);
      float riz, riz_dup;
      pk2_floats(riz2, riz, riz_dup);
      const double2 ui = lds_double2(rq + 32u);
      const double uix = ui.x, uiy = ui.y, uiz = __longlong_as_double((long long)uiz2);
      const int moli = EXCL_MOL ? s_mol[il] : 0;
      const int il_eff = pass0 ? il : -1;
#pragma unroll
      for (int c = 0; c < SLOTS; ++c) {
        const int g = c < ns ? s_gid[c][lane] : -1;
        bool pc = g >= 0;
        if (c < HSL) pc = pc && 32 * c + lane > il_eff;
        if (EXCL_MOL) pc = pc && molr[c] != moli;
        if (pc) {
          float px0, px1, py0, py1, pz0, pz1;
          pk2_floats(xr2[c >> 1], px0, px1);
          pk2_floats(yr2[c >> 1], py0, py1);
          pk2_floats(zr2[c >> 1], pz0, pz1);
          const float px = (c & 1) ? px1 : px0, py = (c & 1) ? py1 : py0, pz = (c & 1) ? pz1 : pz0;
          const float dx = __fadd_rn(px, ria.x), dy = __fadd_rn(py, ria.z), dz = __fadd_rn(pz, riz);
          const float t = __fadd_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))), k.nlo);
          if (__float_as_uint(t) <= k.band_bits) {
            fix_append(a.counters, a.fix_list, a.fix_cap, hs + il, g, 0);          // undecided: nothing was added
          } else if (t < 0.0f) {
            const double d = fma(uix, ux[c], fma(uiy, uy[c], uiz * uz[c]));
            if (fabs(d) < HS_NEARZ) fix_append(a.counters, a.fix_list, a.fix_cap, hs + il, g, 1);   // counted; cos^2 == 0 ?
          }
        }
      }
    }
    // ---- reference side: per-lane partials of the group's references are reduced together
    {
      __syncwarp();
      const int row = lane & (HS_GRP - 1), part = lane / HS_GRP;
      // fixed pairwise tree (deterministic, log-depth dependency chain)
      double pv[HS_GRP];
      int nv = 0;
#pragma unroll
      for (int q = 0; q < HS_GRP; ++q) {
        pv[q] = s_acc[row][part * HS_GRP + q];
        nv += (int)s_cnt[row][part * HS_GRP + q];
      }
#pragma unroll
      for (int w = 1; w < HS_GRP; w <<= 1)
#pragma unroll
        for (int q = 0; q + w < HS_GRP; q += 2 * w) pv[q] += pv[q + w];
      // the 32 / HS_GRP partials of a reference go straight to the integer accumulators (no shuffle stages:
      // their latency would be exposed here, the extra atomics are not)
      MOUSE_ASSERT(row >= nref || (hs + base + row >= 0 && hs + base + row < a.n));
      if (row < nref) hs_flush(a.acc_fix, a.acc_cnt, hs + base + row, pv[0], nv);
      __syncwarp();
    }
  }
}

// Drops the candidates of the landing zone that are farther than the cutoff from the bounding box of the home cell's
// bonds (no reference of the home cell can reach them): s_sel[0..M) = landing-zone index of the candidates kept, in
// order, as BYTE offsets (48 * index); returns M.  bb (shared memory) = home cell centre, box lower corner, box upper corner.  BND: the home cell
// sits on the periodic boundary; the candidates are first moved to their image nearest to the home cell centre
// (written back to the landing zone).  Two candidates per lane and trip (two independent chains).
#define HS_CULL_W 2        // candidates per lane and trip of the cull (independent chains: the trip is latency-bound)
template <bool BND>
__device__ __forceinline__ int hs_cull(uint32_t land, uint32_t sel, const float *bb, const GridDev &g, int lane,
                                       int nland, float thr) {
  const float blx = bb[3], bly = bb[4], blz = bb[5], bhx = bb[6], bhy = bb[7], bhz = bb[8];
  const unsigned lt = (1u << lane) - 1u;
  int M = 0;
  uint32_t pa = land + 48u * (uint32_t)lane;
  for (int t0 = 0; t0 < nland; t0 += 32 * HS_CULL_W, pa += 32u * HS_CULL_W * 48u) {
    float4 p[HS_CULL_W];
#pragma unroll
    for (int j = 0; j < HS_CULL_W; ++j) p[j] = lds_float4(pa + 32u * 48u * j);
    if (BND) {
      const float hx = bb[0], hy = bb[1], hz = bb[2];
#pragma unroll
      for (int j = 0; j < HS_CULL_W; ++j) {
        hs_nearest_image(p[j].x, p[j].y, p[j].z, hx, hy, hz, g);
        sts_float4(pa + 32u * 48u * j, p[j]);
      }
    }
    unsigned bal[HS_CULL_W];
    bool keep[HS_CULL_W];
#pragma unroll
    for (int j = 0; j < HS_CULL_W; ++j) {
      const float ex = fmaxf(fmaxf(__fadd_rn(blx, -p[j].x), __fadd_rn(p[j].x, -bhx)), 0.0f);
      const float ey = fmaxf(fmaxf(__fadd_rn(bly, -p[j].y), __fadd_rn(p[j].y, -bhy)), 0.0f);
      const float ez = fmaxf(fmaxf(__fadd_rn(blz, -p[j].z), __fadd_rn(p[j].z, -bhz)), 0.0f);
      const float d2 = __fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex)));
      keep[j] = (t0 + 32 * j + lane < nland) & (d2 <= thr);
      bal[j] = __ballot_sync(MOUSE_FULL_MASK, keep[j]);
    }
#pragma unroll
    for (int j = 0; j < HS_CULL_W; ++j) {
      if (keep[j]) sts_u16(sel + 2u * (uint32_t)(M + __popc(bal[j] & lt)), 48 * (t0 + 32 * j + lane));
      M += __popc(bal[j]);
    }
  }
  __syncwarp();
  return M;
}

template <int SLOTS, bool EXCL_MOL>
__global__ void __launch_bounds__(32, HS_WARPS) p2_sweep_half_kernel(const SweepArgs a) {
  constexpr int NC = SLOTS * 32;   // register-resident candidates per sub-pass
  constexpr int NPAIR = SLOTS / 2;
  constexpr int HSL = HS_MAXHOME / 32;
  constexpr int LZ = HS_LZ;
  static_assert(LZ % HS_MAXHOME == 0, "a window of references never straddles two chunks");
  static_assert(SLOTS % 2 == 0 && NC >= HS_MAXHOME && LZ >= HS_MAXHOME && LZ % (32 * HS_CULL_W) == 0 && NPAIR <= 6,
                "slots come in pairs; the home cell lives in the first sub-pass of the first chunk");
  __shared__ __align__(128) Rec s_land[LZ + 1];     // TMA landing zone (one chunk of the candidate concatenation);
                                                    // record LZ is the filler of the empty register slots
  __shared__ uint16_t s_sel[LZ + NC];               // byte offset in the landing zone of the candidates that survived
                                                    // the cull, then NC times the filler
  __shared__ HsRef s_ref[HS_MAXHOME + 1];           // the references of this window (+ the far-away sentinel)
  __shared__ int s_mol[EXCL_MOL ? HS_MAXHOME + 2 : 1];   // their molecule codes (only read when pairs inside a molecule are excluded)
  __shared__ double s_acc[HS_GRP][33];              // per-lane partial sums of 8 references (transposed reduce)
  __shared__ uint8_t s_cnt[HS_GRP][36];             // ... and partial counts (<= SLOTS per lane and reference: a byte each)
  __shared__ int s_gid[SLOTS][32];                  // sorted slot of every register-resident candidate (-1: empty)
  __shared__ __align__(16) int s_raw[HS_ITEM_INTS]; // item row in flight (fetch stage 2 -> 3)
  __shared__ int s_next[3][10];                     // range table of the prefetched item (kept out of the registers)
  __shared__ float s_bb[9];                         // home cell centre + bounding box of the prefetched / current item
  __shared__ int s_nexti[6];
  __shared__ __align__(8) uint64_t s_bar;
  const int lane = threadIdx.x;
  HsConst k;
  k.nlo = -a.g.rc2_lo;                                    // guard band [rc2_lo, rc2_hi] around rc^2
  k.band_bits = __float_as_uint((a.g.rc2_hi - a.g.rc2_lo) * 1.0001f);
  k.nearz_thr = __int_as_float(HS_NEARZ_HI);
  const float cull_thr = a.g.rc2_hi * 1.001f;             // box distance test: well outside the guard band

  if (lane == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  for (int q = lane; q < LZ + NC; q += 32) s_sel[q] = (uint16_t)(48 * LZ);
  if (lane == 0) {
    s_land[LZ].p = make_float4(HS_PAD, HS_PAD, HS_PAD, __int_as_float(-1));
    s_land[LZ].u = make_double4(1.0, 1.0, 1.0, __hiloint2double(0, -1));
  }
  __syncwarp();
  uint32_t phase = 0;

  // arm the landing zone with chunk `ch` of item `it`: candidates [ch * LZ, min(T, ch * LZ + LZ))
  auto issue = [&](const HsItem &it, int ch) {
    const int w0 = ch * LZ, w1 = min(it.T, w0 + LZ);
    MOUSE_ASSERT(ch >= 0 && w1 > w0 && w1 - w0 <= LZ);
    // WAR on the landing zone: every lane's shared-memory accesses to the previous contents were issued before the
    // __syncwarp() that precedes this call; the proxy fence orders them against the async-proxy writes
    fence_proxy_async();
    if (lane == 0) mbar_arrive_expect_tx(&s_bar, (uint32_t)(w1 - w0) * 48u);
    __syncwarp();
    const int a0 = max(it.rt, w0), b0 = min(it.rt + it.rlen, w1);
    if (b0 > a0) {
      const int src = it.rsrc + (a0 - it.rt), cnt = b0 - a0, dst = a0 - w0;
      MOUSE_ASSERT(dst >= 0 && dst + cnt <= LZ && src >= 0 && src + cnt <= a.n);
      tma_load_1d(&s_land[dst], a.R + src, (uint32_t)cnt * 48u, &s_bar);
    }
  };

  // Three-stage fetch pipeline, one stage per loop iteration so that no global latency is exposed:
  //   stage 1  ticket  = atomicAdd(home-cell counter)           (tkP: issued one iteration before it is used)
  //   stage 2  item row of the ticket's home cell -> s_raw      (tkA: cp.async issued one iteration before stage 3)
  //   stage 3  finalize the row, arm the landing zone (TMA)     (one iteration before the data is consumed)
  // item rows of the frame: the sum of the class histogram the scan left in counters[16..23]
  const int nrows = warp_sum(lane < 8 ? (int)a.counters[16 + lane] : 0);
  int tkA, tkP = 0;
  {
    int t0 = 0;
    if (lane == 0) {
      t0 = (int)atomicAdd((unsigned int *)&a.counters[C_TICKET], 1u);    // low word of the 64-bit counter
      tkP = (int)atomicAdd((unsigned int *)&a.counters[C_TICKET], 1u);
    }
    tkA = __shfl_sync(MOUSE_FULL_MASK, t0, 0);
    if (tkA < nrows) hs_raw_issue(a.items, lane, tkA, s_raw);
  }
  // next non-empty, regular home cell (its centre and bounding box go to s_bb); false when the grid is exhausted
  auto advance = [&](HsItem &it) -> bool {
    for (;;) {
      if (tkA >= nrows) return false;
      const bool nonempty = hs_finalize(a.g, lane, tkA, s_raw, it);
      const float bbv = lane < 6 ? __int_as_float(s_raw[32 + lane]) : 0.0f;
      __syncwarp();
      tkA = __shfl_sync(MOUSE_FULL_MASK, tkP, 0);
      if (lane == 0) tkP = (int)atomicAdd((unsigned int *)&a.counters[C_TICKET], 1u);
      if (tkA < nrows) hs_raw_issue(a.items, lane, tkA, s_raw);
      if (!nonempty) continue;
      if (it.T > HS_MAXT) {
        if (it.win == 0)       // (once per cell, not once per window row)
          hs_oversize(a.R, a.counters, a.fix_list, a.fix_cap, a.same_molecule, lane, it.hs, it.nhome, it.rsrc, it.rlen, it.rt);
        continue;
      }
      if (lane < 6) s_bb[3 + lane] = bbv;
      if (lane == 0) {
        s_bb[0] = it.hx;
        s_bb[1] = it.hy;
        s_bb[2] = it.hz;
      }
      return true;
    }
  };

  // A home cell with more than HS_MAXHOME bonds is taken in windows of HS_MAXHOME references [rb, rb + HS_MAXHOME),
  // one ITEM ROW per window (cell_items_kernel): each window is a pass over the candidate concatenation of its own
  // that starts in the chunk holding its references and skips the survivors in front of the window (home-cell bonds
  // j < rb <= i: those pairs belong to the earlier windows).  The windows of a crowded cell are separate tickets, so
  // several warps share the cell (one warp sweeping a 65-bond cell alone takes as long as a whole 250k-bond frame).
  HsItem cur;
  bool have = advance(cur);
  int ch = have ? (cur.win * HS_MAXHOME) / LZ : 0, sp = 0, M = 0, rb = have ? cur.win * HS_MAXHOME : 0;
  bool fresh = true;
  if (have) issue(cur, ch);

  while (have) {
    if (fresh) {
      // ---- a new chunk has landed: drop the candidates no reference of the home cell can reach
      const int nland = min(cur.T - ch * LZ, LZ);
      MOUSE_ASSERT(nland > 0 && cur.hs >= 0 && cur.nhome > 0 && rb >= 0 && rb < cur.nhome && cur.hs + cur.nhome <= a.n);
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
      M = cur.boundary ? hs_cull<true>(smem_addr(s_land), smem_addr(s_sel), s_bb, a.g, lane, nland, cull_thr)
                       : hs_cull<false>(smem_addr(s_land), smem_addr(s_sel), s_bb, a.g, lane, nland, cull_thr);
      MOUSE_ASSERT(M >= 0 && M <= nland);
      // the NC entries behind the survivors point at the filler record
#pragma unroll
      for (int c = 0; c < SLOTS; ++c) s_sel[M + 32 * c + lane] = (uint16_t)(48 * LZ);
      __syncwarp();
      sp = 0;
    }
    // ---- landing zone -> registers: candidates [k0, k0 + NC) of the survivors (first chunk: behind the window start)
    const int k0 = sp * NC + (ch == rb / LZ ? rb % LZ : 0);
    const int nvalid = max(min(M - k0, NC), 0);
    const int npp = (nvalid + 63) >> 6;             // slot pairs in use
    const bool last = k0 + NC >= M;                 // last sub-pass of this chunk
    const bool pass0 = (ch == rb / LZ) & (sp == 0);   // the home cell leads the concatenation and is never dropped
    const int hs = cur.hs + rb, nhome = min(cur.nhome - rb, HS_MAXHOME);   // this window's references
    pk2 xr2[NPAIR], yr2[NPAIR], zr2[NPAIR];
    double ux[SLOTS], uy[SLOTS], uz[SLOTS];
    double accj[SLOTS];
    int cntj[SLOTS], molr[SLOTS];
    {
      const uint32_t land = smem_addr(s_land), selp = smem_addr(s_sel) + 2u * (uint32_t)(k0 + lane);
      uint32_t off[SLOTS];
#pragma unroll
      for (int c = 0; c < SLOTS; ++c) off[c] = land + lds_u16(selp + 64u * c);
#pragma unroll
      for (int pp = 0; pp < NPAIR; ++pp) {
        float4 p[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = 2 * pp + h;
          p[h] = lds_float4(off[c]);
          const double2 u01 = lds_double2(off[c] + 16u), u2w = lds_double2(off[c] + 32u);
          ux[c] = u01.x;
          uy[c] = u01.y;
          uz[c] = u2w.x;
          s_gid[c][lane] = __double2loint(u2w.y);      // the filler record carries -1
          molr[c] = __float_as_int(p[h].w);
          accj[c] = 0.0;
          cntj[c] = 0;
          if (c < HSL && pass0 && 32 * c + lane < nhome) {   // survivors [0, nhome) of the cell are its own bonds, in order
            HsRef rr;
            rr.a = make_float4(-p[h].x, -p[h].x, -p[h].y, -p[h].y);
            rr.b = make_float2(-p[h].z, -p[h].z);
            rr.uz = u2w.x;
            rr.u = u01;
            s_ref[32 * c + lane] = rr;
            if (EXCL_MOL) s_mol[32 * c + lane] = __float_as_int(p[h].w);
          }
        }
        // (the packed add of -0.0 makes each pair the RESULT of an instruction: the compiler then keeps the pair in
        // an aligned register pair instead of re-pairing two scalar registers inside the reference loop)
        const pk2 nz2 = pk2_make(-0.0f, -0.0f);
        xr2[pp] = pk2_add(pk2_make(p[0].x, p[1].x), nz2);
        yr2[pp] = pk2_add(pk2_make(p[0].y, p[1].y), nz2);
        zr2[pp] = pk2_add(pk2_make(p[0].z, p[1].z), nz2);
      }
      if (pass0 && lane == 0) {      // sentinel behind the window: the partner of the last reference of an odd tail
        HsRef rr;
        rr.a = make_float4(HS_PAD, HS_PAD, HS_PAD, HS_PAD);       // at -HS_PAD: far from the bonds AND from the filler
        rr.b = make_float2(HS_PAD, HS_PAD);
        rr.uz = 0.0;
        rr.u = make_double2(1.0, 0.0);
        s_ref[nhome] = rr;
        if (EXCL_MOL) s_mol[nhome] = -1;
      }
    }
    __syncwarp();

    // ---- after the last sub-pass's register load the landing zone is free: re-arm it with the next chunk of
    // this home cell, or with the first chunk of the next one (the copy then overlaps the arithmetic below)
    int nch = ch + 1, nrb = rb;
    if (last) {
      HsItem nxt = cur;
      if (nch * LZ >= cur.T) {               // this window is complete: next item row
        have = advance(nxt);
        nrb = have ? nxt.win * HS_MAXHOME : 0;
        nch = nrb / LZ;                      // the chunk the window starts in
      }
      if (have) {
        issue(nxt, nch);
        if (lane < 10) {
          s_next[0][lane] = nxt.rsrc;
          s_next[1][lane] = nxt.rlen;
          s_next[2][lane] = nxt.rt;
        }
        if (lane == 0) {
          s_nexti[0] = nxt.win;
          s_nexti[1] = nxt.hs;
          s_nexti[2] = nxt.nhome;
          s_nexti[3] = nxt.T;
          s_nexti[4] = nxt.boundary;
        }
      }
    }

    // ---- the references of the home cell against the register-resident candidates
    if (nvalid > 0)      // (a later chunk may have no survivors)
      hs_ref_loop<SLOTS, EXCL_MOL>(a, k, lane, pass0, (nvalid + 31) >> 5, hs, nhome, xr2, yr2, zr2, ux, uy, uz, accj, cntj, molr,
                                   smem_addr(s_ref), s_mol, s_acc, s_cnt, s_gid);
    // ---- candidate side: one pair of integer atomics per candidate that was hit
#pragma unroll
    for (int c = 0; c < SLOTS; ++c) {      // (slots that were not in use and fillers have a zero count: predicated off)
      const int g = s_gid[c][lane];
      MOUSE_ASSERT(g < a.n && (g >= 0 || cntj[c] == 0));
      hs_flush(a.acc_fix, a.acc_cnt, g < 0 ? 0 : g, accj[c], cntj[c]);
    }
    if (!last) {
      sp += 1;
      fresh = false;
      __syncwarp();
      continue;
    }
    if (!have) break;
    __syncwarp();
    cur.rsrc = s_next[0][lane < 10 ? lane : 0];
    cur.rlen = lane < 10 ? s_next[1][lane] : 0;
    cur.rt = s_next[2][lane < 10 ? lane : 0];
    cur.win = s_nexti[0];
    cur.hs = s_nexti[1];
    cur.nhome = s_nexti[2];
    cur.T = s_nexti[3];
    cur.boundary = s_nexti[4];
    ch = nch;
    rb = nrb;
    fresh = true;
    __syncwarp();
  }
}

// ------------------------------------------------------------------ exact all-FP64 sweep
template <bool LISTS>
__device__ __forceinline__ void p2_sweep_exact_body(const SweepArgs &a) {
  const int n_list = a.cell_start[a.g.ncell];
  const size_t n = (size_t)a.n;
  const int nx = a.g.nc[0], ny = a.g.nc[1], nz = a.g.nc[2];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_list; s += gridDim.x * blockDim.x) {
    if (!(a.flag[s] & 1)) continue;
    const int bi = a.idx[s];
    const double cix = a.ctr[bi], ciy = a.ctr[n + bi], ciz = a.ctr[2 * n + bi];
    const double bix = a.vec[bi], biy = a.vec[n + bi], biz = a.vec[2 * n + bi];
    const double nr = __dadd_rn(__dadd_rn(__dmul_rn(bix, bix), __dmul_rn(biy, biy)), __dmul_rn(biz, biz));
    const int moli = __float_as_int(a.R[s].p.w);
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
            if (!a.same_molecule && __float_as_int(a.R[t].p.w) == moli) continue;
            const int bj = a.idx[t];
            if (!a.no_cutoff) {
              double rsq = mouse_rsq_exact_g(cix, ciy, ciz, a.ctr[bj], a.ctr[n + bj], a.ctr[2 * n + bj], a.g.box,
                                             a.g.tilt, a.g.tri);
              if (!(rsq <= a.g.rc2)) continue;
            }
            const double bjx = a.vec[bj], bjy = a.vec[n + bj], bjz = a.vec[2 * n + bj];
            const double dot = mouse_dot_exact(bix, biy, biz, bjx, bjy, bjz);
            const double nj = __dadd_rn(__dadd_rn(__dmul_rn(bjx, bjx), __dmul_rn(bjy, bjy)), __dmul_rn(bjz, bjz));
            const double cs = __ddiv_rn(__ddiv_rn(__dmul_rn(dot, dot), nr), nj);
            if (cs != 0.0) {
              if (LISTS && row + cnt < a.list_cap) a.nbr_idx[row + cnt] = bj;
              sum += cs;
              cnt += 1;
            }
          }
        }
      }
    }
    if (!LISTS) {
      const long long nzero = a.counters[C_NZERO];
      if (nzero) {
        sum = __longlong_as_double(0x7ff8000000000000LL);
        cnt += (int)nzero;
      }
      a.out_sum[bi] = sum;
      a.out_cnt[bi] = cnt;
    }
  }
}

template <bool LISTS>
__global__ void __launch_bounds__(128) p2_sweep_exact_kernel(const SweepArgs a) {
  p2_sweep_exact_body<LISTS>(a);
}

// Exact decision of the listed pairs (one thread per pair): kind 0 = nothing was added by the fast sweep, add the
// reference's cos^2 to both partners if the pair is counted; kind 1 = added as a hit with a near-zero dot product,
// remove it from both counts if the reference's cos^2 is exactly 0.0 (masked, :113).
// If the pair list overflowed (pathological inputs: e.g. a lattice with most pairs exactly at the cutoff) the same
// launch recomputes the whole frame with the all-FP64 sweep instead: results stay exact, only slower.
// (64-thread blocks: they fit next to the persistent sweep of another frame in flight.)
#define HS_FIXUP_THREADS 64
__global__ void __launch_bounds__(HS_FIXUP_THREADS) p2_fixup_kernel(const SweepArgs a) {
  if (((volatile int64_t *)a.counters)[C_OVERFLOW]) {
    p2_sweep_exact_body<false>(a);
    return;
  }
  const long long nfix = ((volatile int64_t *)a.counters)[C_NFIX];
  const size_t n = (size_t)a.n;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nfix; e += (long long)gridDim.x * blockDim.x) {
    const int2 pr = a.fix_list[e];
    const int si = pr.x & 0x7fffffff, sj = pr.y, kind = (unsigned)pr.x >> 31;
    MOUSE_ASSERT(si < a.n && sj >= 0 && sj < a.n && si != sj);
    const double cs = exact_pair(a.idx, a.ctr, a.vec, n, a.g, si, sj);
    if (kind == 0) {
      if (cs != 0.0) {
        const unsigned long long f = (unsigned long long)__double2ll_rn(cs * HS_FIX_SCALE);
        atomicAdd((unsigned long long *)&a.acc_fix[si], f);
        atomicAdd((unsigned long long *)&a.acc_fix[sj], f);
        atomicAdd(&a.acc_cnt[si], 1);
        atomicAdd(&a.acc_cnt[sj], 1);
      }
    } else if (cs == 0.0) {
      atomicAdd(&a.acc_cnt[si], -1);
      atomicAdd(&a.acc_cnt[sj], -1);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[C_NFIX_DONE] = nfix;
}

// accumulators (sorted slots, fixed point) -> out_sum / out_cnt (original bond order).  The accumulators are
// handed back zeroed: the cell build zeroes them once, every sweep leaves them as it found them.
__global__ void __launch_bounds__(256) p2_finish_kernel(const SweepArgs a) {
  const bool overflow = ((volatile int64_t *)a.counters)[C_OVERFLOW] != 0;   // the exact sweep writes the outputs
  const int n_list = a.cell_start[a.g.ncell];
  const long long nzero = a.counters[C_NZERO];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_list; s += gridDim.x * blockDim.x) {
    const long long fix = a.acc_fix[s];
    int cnt = a.acc_cnt[s];
    a.acc_fix[s] = 0;
    a.acc_cnt[s] = 0;
    if (overflow || !(a.flag[s] & 1)) continue;     // not a reference: out stays 0 / 0
    const int o = a.idx[s];
    double sum = (double)fix * HS_FIX_INV;
    // zero-length bonds in the neighbour set turn every mean into NaN and add to every count
    // (0/0 in ordering_functions.py:112)
    if (nzero) {
      sum = __longlong_as_double(0x7ff8000000000000LL);
      cnt += (int)nzero;
    }
    a.out_sum[o] = sum;
    a.out_cnt[o] = cnt;
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
// Immutable per-device facts (SM count, resident sweep CTAs per SM), looked up once per device ordinal.
struct DevFacts {
  int sms;
  int per_sm[8];   // by kernel variant
};
static DevFacts g_dev[64];

static DevFacts *dev_facts() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  DevFacts *f = &g_dev[dev];
  if (f->sms == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    f->sms = n > 0 ? n : 148;
  }
  return f;
}

int mouse_num_sms() { return dev_facts()->sms; }

template <int SLOTS, bool EXCL>
static void launch_half_one(const SweepArgs &a, int variant, int warps_override, cudaStream_t st) {
  DevFacts *f = dev_facts();
  if (f->per_sm[variant] == 0) {
    int b = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, p2_sweep_half_kernel<SLOTS, EXCL>, 32, 0);
    f->per_sm[variant] = b < 1 ? 1 : b;
  }
  // persistent warps: every resident warp (CTA) pulls home cells from the ticket counter
  int per_sm = f->per_sm[variant];
  if (warps_override > 0 && warps_override < per_sm) per_sm = warps_override;
  int blocks = f->sms * per_sm;
  if (blocks > a.g.ncell) blocks = a.g.ncell;     // (at least one row per cell)
  p2_sweep_half_kernel<SLOTS, EXCL><<<blocks, 32, 0, st>>>(a);
}

template <int SLOTS>
static void launch_half(const SweepArgs &a, bool excl_mol, int variant, int warps_override, cudaStream_t st) {
  if (excl_mol) launch_half_one<SLOTS, true>(a, 2 * variant + 1, warps_override, st);
  else launch_half_one<SLOTS, false>(a, 2 * variant, warps_override, st);
}

// The FP32-filter half-shell sweep applies when the grid allows it (>= 3 cells per axis, small guard band).  Crowded
// cells are taken in windows of references, one item row each, so the fast path stays ahead of the all-FP64 sweep far
// beyond the landing zone (250k bonds at rc = 7 sigma, 343 per cell: 2.4 ms against 15.4 ms; rc = 9 sigma, 729 per
// cell: 6.8 against 28.9 ms); the limit is the candidate count a cell's fixed-point accumulators are sized for
// (HS_MAXT = 16 384 per home cell): frames with more than 512 items per cell on average go to the all-FP64 sweep.
int mouse_p2_uses_fast(const mouse_grid_t *grid, int64_t n, uint32_t flags, bool lists) {
  if (!grid->fast || lists || (flags & (MOUSE_FORCE_EXACT | MOUSE_NO_CUTOFF))) return 0;
  return (flags & MOUSE_FORCE_FAST) || (double)n <= 512.0 * (double)grid->ncell;
}

// frame_mode 0: stand-alone sweep (outputs zeroed here, accumulators converted by p2_finish_kernel);
// frame_mode 1: called from mouse_p2_frame - the outputs were initialised by the build and the caller converts the
// accumulators in the fused finish + reduce kernel.  *used_fast tells the caller which path ran.
int mouse_launch_p2_sweep(const double *vec, const double *ctr, int64_t n, const mouse_grid_t *grid, uint32_t flags,
                          const Workspace &w, double *out_sum, int32_t *out_cnt, const int64_t *row_ptr,
                          int32_t *nbr_idx, int64_t list_cap, int frame_mode, int *used_fast, cudaStream_t st) {
  SweepArgs a;
  a.g = mouse_grid_dev(grid);
  a.R = w.rec_sorted;
  a.idx = w.idx_sorted;
  a.cell_start = w.cell_start;
  a.items = w.items;
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
  a.acc_fix = w.acc_fix;
  a.acc_cnt = w.acc_cnt;
  a.same_molecule = (flags & MOUSE_SAME_MOLECULE) ? 1 : 0;
  a.no_cutoff = (flags & MOUSE_NO_CUTOFF) ? 1 : 0;
  a.list_cap = list_cap;
  a.fix_list = (int2 *)w.rec_tmp;    // the original-order payload is dead once the cell list is built
  a.fix_cap = 6LL * (n > 0 ? n : 1);
  const bool lists = nbr_idx != nullptr;
  const bool fast = mouse_p2_uses_fast(grid, n, flags, lists) != 0;
  if (used_fast) *used_fast = fast ? 1 : 0;
  if (n == 0) return MOUSE_OK;
  if (!frame_mode) {
    if (!lists) MOUSE_CUDA_CHECK(cudaMemsetAsync(out_sum, 0, sizeof(double) * (size_t)n, st));
    MOUSE_CUDA_CHECK(cudaMemsetAsync(out_cnt, 0, sizeof(int32_t) * (size_t)n, st));
    // counters 1..4: diagnostics, home-cell ticket, pair list length, overflow flag (the reduce kernel resets
    // the ticket again before using it); a frame build has just cleared all of them
    MOUSE_CUDA_CHECK(cudaMemsetAsync(w.counters + 1, 0, 4 * sizeof(int64_t), st));
    if (grid->fast) {   // a fused frame on the same cell list leaves its sums in the accumulators
      MOUSE_CUDA_CHECK(cudaMemsetAsync(w.acc_fix, 0, sizeof(long long) * (size_t)n, st));
      MOUSE_CUDA_CHECK(cudaMemsetAsync(w.acc_cnt, 0, sizeof(int32_t) * (size_t)n, st));
    }
  }
  const int sms = mouse_num_sms();
  if (fast) {
    // SLOTS * 32 register-resident candidates per sub-pass, sized from the mean of the 14 half-shell cells (about a
    // third of them is dropped by the bounding-box cull)
    double mean = 14.0 * (double)n / (double)grid->ncell;
    int slots = mean <= 150.0 ? 4 : 6;
    const int slots_override = (int)((flags >> 8) & 0xffu), warps_override = (int)((flags >> 16) & 0xffu);
    if (slots_override) slots = slots_override;
    const bool excl = !a.same_molecule;
    switch (slots) {
      case 4: launch_half<4>(a, excl, 0, warps_override, st); break;
      case 6: launch_half<6>(a, excl, 1, warps_override, st); break;
      default: launch_half<8>(a, excl, 2, warps_override, st); break;
    }
    MOUSE_LAUNCH_CHECK("p2_sweep_half_kernel");
    p2_fixup_kernel<<<sms * 8, HS_FIXUP_THREADS, 0, st>>>(a);     // (or the exact redo of the frame after an overflow)
    MOUSE_LAUNCH_CHECK("p2_fixup_kernel");
    if (!frame_mode) {
      p2_finish_kernel<<<sms * 8, 256, 0, st>>>(a);
      MOUSE_LAUNCH_CHECK("p2_finish_kernel");
    }
  } else {
    int blocks = (int)((n + 127) / 128);
    if (blocks > sms * 32) blocks = sms * 32;
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
