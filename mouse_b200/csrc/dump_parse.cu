// Device-side parser of the ATOMS block of a LAMMPS dump frame (file ingestion row).
// The reference reads the block with one float() per token (lammpack_types.py:377-395) and applies
// `lo + (hi - lo) * xs`; the host parser of api.cu (mouse_parse_dump_atoms) does the same in C.  Here the raw text of
// the block is shipped to the GPU as it lies in the file (no host parsing at all: the host only moves bytes from the
// page cache into a pinned buffer) and turned into ids / coordinates / image flags by four small kernels:
//   dump_count_lines_kernel   non-blank lines per 4 KB tile
//   dump_scan_tiles_kernel    exclusive scan of the tile counts (one block), total -> status[0]
//   dump_fill_lines_kernel    byte offset of every non-blank line, in file order
//   dump_parse_rows_kernel    one thread per line: tokens -> numbers, the reference's coordinate transform
// Numbers are correctly rounded exactly as on the host: <= 15 significant digits and a decimal exponent within +-22
// = integer mantissa times / divided by an exact power of ten, ONE IEEE rounding (__dmul_rn / __ddiv_rn).  Everything
// else (longer mantissas, nan / inf, malformed tokens, short or long lines) is only COUNTED: the caller then parses
// that frame with the host parser, which is authoritative for errors and for the strtod cases.
#include "common.cuh"

#define DP_TILE_THREADS 256
#define DP_BYTES_PER_THREAD 16
#define DP_TILE_BYTES (DP_TILE_THREADS * DP_BYTES_PER_THREAD)

struct DumpCols {
  int ncol, col_id, mode, has_img;
  int col_xyz[3], col_img[3];
  double lo[3], span[3];
};

__device__ __forceinline__ bool dp_ws(unsigned c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }

// byte k opens a line (k == 0 or behind a newline): does the line hold anything but whitespace?
__device__ __forceinline__ bool dp_line_not_blank(const unsigned char *t, int64_t len, int64_t k) {
  for (; k < len; ++k) {
    const unsigned c = t[k];
    if (c == '\n') return false;
    if (!dp_ws(c)) return true;
  }
  return false;
}

// non-blank lines that start inside this thread's 16 bytes; bit i of the result = byte k0 + i opens one
__device__ __forceinline__ unsigned dp_thread_starts(const unsigned char *t, int64_t len, int64_t k0) {
  if (k0 >= len) return 0u;
  unsigned char c[DP_BYTES_PER_THREAD];
  if (k0 + DP_BYTES_PER_THREAD <= len) {
    const uint4 w = *reinterpret_cast<const uint4 *>(t + k0);      // (t is 16-byte aligned, k0 a multiple of 16)
    memcpy(c, &w, sizeof(w));
  } else {
#pragma unroll
    for (int i = 0; i < DP_BYTES_PER_THREAD; ++i) c[i] = k0 + i < len ? t[k0 + i] : (unsigned char)'\n';
  }
  unsigned prev = k0 > 0 ? t[k0 - 1] : (unsigned)'\n';
  unsigned m = 0;
#pragma unroll
  for (int i = 0; i < DP_BYTES_PER_THREAD; ++i) {
    if (prev == '\n' && k0 + i < len) {
      const bool text_here = !dp_ws(c[i]);      // the usual case: no leading whitespace, no look-ahead
      if (text_here || (c[i] != '\n' && dp_line_not_blank(t, len, k0 + i))) m |= 1u << i;
    }
    prev = c[i];
  }
  return m;
}

__global__ void __launch_bounds__(DP_TILE_THREADS) dump_count_lines_kernel(const unsigned char *t, int64_t len,
                                                                         unsigned *tile_count) {
  const int64_t k0 = ((int64_t)blockIdx.x * DP_TILE_THREADS + threadIdx.x) * DP_BYTES_PER_THREAD;
  const int n = __popc(dp_thread_starts(t, len, k0));
  __shared__ int s_w[DP_TILE_THREADS / 32];
  const int w = __reduce_add_sync(MOUSE_FULL_MASK, n);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < DP_TILE_THREADS / 32; ++i) s += s_w[i];
    tile_count[blockIdx.x] = (unsigned)s;
  }
}

// exclusive scan of the tile counts in place (one block); status[0] = lines in the block of text, [1] = no error yet,
// [2] = [3] = 0
__global__ void __launch_bounds__(1024) dump_scan_tiles_kernel(unsigned *tile_count, int ntiles, long long *status) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_carry;
  if (threadIdx.x == 0) s_carry = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ntiles; base += 1024) {
    const int k = base + threadIdx.x;
    const unsigned v = k < ntiles ? tile_count[k] : 0u;
    unsigned x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned y = __shfl_up_sync(MOUSE_FULL_MASK, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
      unsigned ws = s_w[lane], xs = ws;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(MOUSE_FULL_MASK, xs, d);
        if (lane >= d) xs += y;
      }
      s_w[lane] = xs - ws;      // exclusive over the warps
    }
    __syncthreads();
    const unsigned carry = s_carry;
    if (k < ntiles) tile_count[k] = carry + s_w[wid] + x - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + s_w[wid] + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    status[0] = (long long)s_carry;
    status[1] = 0x7fffffffffffffffLL;
    status[2] = 0;
    status[3] = 0;
  }
}

__global__ void __launch_bounds__(DP_TILE_THREADS) dump_fill_lines_kernel(const unsigned char *t, int64_t len,
                                                                        const unsigned *tile_base, int64_t nrows,
                                                                        unsigned *line_off) {
  const int64_t k0 = ((int64_t)blockIdx.x * DP_TILE_THREADS + threadIdx.x) * DP_BYTES_PER_THREAD;
  unsigned m = dp_thread_starts(t, len, k0);
  const int n = __popc(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __shared__ int s_w[DP_TILE_THREADS / 32];
  int x = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(MOUSE_FULL_MASK, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) s_w[wid] = x;
  __syncthreads();
  int before = 0;
#pragma unroll
  for (int i = 0; i < DP_TILE_THREADS / 32; ++i) before += i < wid ? s_w[i] : 0;
  int64_t row = (int64_t)tile_base[blockIdx.x] + before + x - n;
  while (m) {
    const int i = __ffs((int)m) - 1;
    m &= m - 1u;
    if (row < nrows) line_off[row] = (unsigned)(k0 + i);      // (more lines than expected: counted, never written)
    ++row;
  }
}

__device__ const double dp_p10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// One decimal token at t[s..) of a line that ends at `end` (newline or end of text): the grammar and the arithmetic of
// mouse_parse_token's fast path (api.cu).  Returns the index behind the token; ok = false when the token needs the host.
__device__ __forceinline__ int64_t dp_token(const unsigned char *t, int64_t s, int64_t end, double &value, bool &ok) {
  bool neg = false;
  if (s < end && (t[s] == '-' || t[s] == '+')) neg = t[s++] == '-';
  unsigned long long m = 0;
  int digits = 0, frac = 0;
  bool any = false, lead = true;
  while (s < end) {
    const unsigned c = t[s];
    if (c < '0' || c > '9') break;
    any = true;
    if (!(lead && c == '0')) {
      lead = false;
      if (digits < 19) m = m * 10ull + (unsigned long long)(c - '0');
      ++digits;
    }
    ++s;
  }
  if (s < end && t[s] == '.') {
    ++s;
    while (s < end) {
      const unsigned c = t[s];
      if (c < '0' || c > '9') break;
      any = true;
      if (!(lead && c == '0')) {
        lead = false;
        if (digits < 19) m = m * 10ull + (unsigned long long)(c - '0');
        ++digits;
      }
      ++frac;
      ++s;
    }
  }
  int e10 = 0;
  bool fast = any;
  if (any && s < end && (t[s] == 'e' || t[s] == 'E')) {
    int64_t q = s + 1;
    bool eneg = false;
    if (q < end && (t[q] == '-' || t[q] == '+')) eneg = t[q++] == '-';
    if (q < end && t[q] >= '0' && t[q] <= '9') {
      int e = 0;
      while (q < end && t[q] >= '0' && t[q] <= '9') {
        if (e < 10000) e = e * 10 + (int)(t[q] - '0');
        ++q;
      }
      e10 = eneg ? -e : e;
      s = q;
    } else {
      fast = false;
    }
  }
  const bool token_ends = s >= end || dp_ws(t[s]);
  const int e = e10 - frac;
  ok = fast && token_ends && digits <= 15 && e >= -22 && e <= 22;
  if (ok) {
    double v = __ull2double_rn(m);                                            // exact: m < 10^15 < 2^53
    v = e >= 0 ? __dmul_rn(v, dp_p10[e]) : __ddiv_rn(v, dp_p10[-e]);          // one correctly rounded operation
    value = neg ? -v : v;
  } else {
    while (s < end && !dp_ws(t[s])) ++s;
  }
  return s;
}

__global__ void __launch_bounds__(128) dump_parse_rows_kernel(const unsigned char *t, int64_t len, const unsigned *line_off,
                                                              int64_t nrows, const DumpCols c,
                                                              const int32_t *row_of_line, const int64_t *expect_ids,
                                                              int64_t *ids, double *xyz, int32_t *img,
                                                              long long *status) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows || r >= status[0]) return;      // (fewer lines than expected: the caller sees status[0] != nrows)
  const int64_t start = line_off[r];
  MOUSE_ASSERT(start >= 0 && start < len);
  int64_t s = start, end = start;
  while (end < len && t[end] != '\n') ++end;      // tokens never cross the end of the line
  int col = 0, slow = 0;
  bool bad = false;
  double px[3] = {0.0, 0.0, 0.0};
  int iv[3] = {0, 0, 0};
  long long id = 0;
  for (;;) {
    while (s < end && dp_ws(t[s])) ++s;
    if (s >= end) break;
    double v = 0.0;
    bool ok;
    s = dp_token(t, s, end, v, ok);
    if (col >= c.ncol) {
      bad = true;
      break;
    }
    if (!ok) {
      ++slow;
    } else {
      if (col == c.col_id) id = __double2ll_rz(v);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        if (col == c.col_xyz[a]) {
          // mode 0 "scaled": lo + (hi - lo) * x; 1 "cut": lo + x; 2 "uncut": x - un-fused, like the NumPy expressions
          const double p = c.mode == 0 ? __dmul_rn(c.span[a], v) : v;
          px[a] = c.mode == 2 ? v : __dadd_rn(c.lo[a], p);
        }
        if (c.has_img && col == c.col_img[a]) iv[a] = (int)v;
      }
    }
    ++col;
  }
  if (bad || col != c.ncol) {
    atomicMin(&status[1], (long long)start);
    return;
  }
  if (slow) {
    atomicAdd((unsigned long long *)&status[2], (unsigned long long)slow);
    return;
  }
  ids[r] = id;
  if (expect_ids && expect_ids[r] != id) atomicAdd((unsigned long long *)&status[3], 1ull);
  const int64_t d = row_of_line ? (int64_t)row_of_line[r] : r;
  MOUSE_ASSERT(d >= 0 && d < nrows);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    xyz[3 * d + a] = px[a];
    if (c.has_img) img[3 * d + a] = iv[a];
  }
}

extern "C" size_t mouse_parse_dump_ws_bytes(int64_t len, int64_t nrows) {
  if (len < 0) len = 0;
  if (nrows < 0) nrows = 0;
  const size_t ntiles = (size_t)(len + DP_TILE_BYTES - 1) / DP_TILE_BYTES + 1;
  return ((ntiles * 4 + 255) / 256) * 256 + (((size_t)nrows * 4 + 255) / 256) * 256 + 256;
}

extern "C" int mouse_parse_dump_atoms_dev(const char *text, int64_t len, int64_t nrows, int32_t ncol, int32_t col_id,
                                          const int32_t *col_xyz, const int32_t *col_img, int32_t mode,
                                          const double *lo, const double *hi, const int32_t *row_of_line,
                                          const int64_t *expect_ids, int64_t *ids, double *xyz, int32_t *img, void *ws,
                                          size_t ws_bytes, int64_t *status, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!text || len < 0 || len >= ((int64_t)1 << 31) || nrows < 0 || ncol < 1 || col_id < 0 || col_id >= ncol ||
      !col_xyz || mode < 0 || mode > 2 || !lo || !hi || !ids || !xyz || !ws || !status || (col_img && !img)) {
    mouse_set_error("mouse_parse_dump_atoms_dev: bad argument");
    return MOUSE_EINVAL;
  }
  if (((uintptr_t)text & 15u) != 0) {
    mouse_set_error("mouse_parse_dump_atoms_dev: the text must be 16-byte aligned");
    return MOUSE_EINVAL;
  }
  DumpCols c;
  c.ncol = ncol;
  c.col_id = col_id;
  c.mode = mode;
  c.has_img = col_img ? 1 : 0;
  for (int a = 0; a < 3; ++a) {
    if (col_xyz[a] < 0 || col_xyz[a] >= ncol || (col_img && (col_img[a] < 0 || col_img[a] >= ncol))) {
      mouse_set_error("mouse_parse_dump_atoms_dev: column index out of range");
      return MOUSE_EINVAL;
    }
    c.col_xyz[a] = col_xyz[a];
    c.col_img[a] = col_img ? col_img[a] : -1;
    c.lo[a] = lo[a];
    c.span[a] = hi[a] - lo[a];
  }
  if (ws_bytes < mouse_parse_dump_ws_bytes(len, nrows)) {
    mouse_set_error("mouse_parse_dump_atoms_dev: workspace of %zu bytes, %zu needed", ws_bytes,
                    mouse_parse_dump_ws_bytes(len, nrows));
    return MOUSE_EWORKSPACE;
  }
  const int ntiles = (int)((len + DP_TILE_BYTES - 1) / DP_TILE_BYTES);
  unsigned *tile_count = (unsigned *)ws;
  unsigned *line_off = (unsigned *)((char *)ws + (((size_t)(ntiles + 1) * 4 + 255) / 256) * 256);
  const unsigned char *t = (const unsigned char *)text;
  if (ntiles > 0) {
    dump_count_lines_kernel<<<ntiles, DP_TILE_THREADS, 0, st>>>(t, len, tile_count);
    MOUSE_LAUNCH_CHECK("dump_count_lines_kernel");
  }
  dump_scan_tiles_kernel<<<1, 1024, 0, st>>>(tile_count, ntiles, (long long *)status);
  MOUSE_LAUNCH_CHECK("dump_scan_tiles_kernel");
  if (ntiles > 0 && nrows > 0) {
    dump_fill_lines_kernel<<<ntiles, DP_TILE_THREADS, 0, st>>>(t, len, tile_count, nrows, line_off);
    MOUSE_LAUNCH_CHECK("dump_fill_lines_kernel");
    // (a block with fewer lines than nrows leaves line_off short: the caller compares status[0] with nrows BEFORE it
    // trusts any output, and the row kernel below only runs over the lines that exist)
    dump_parse_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(t, len, line_off, nrows, c, row_of_line,
                                                                            expect_ids, ids, xyz, img,
                                                                            (long long *)status);
    MOUSE_LAUNCH_CHECK("dump_parse_rows_kernel");
  }
  return MOUSE_OK;
}
