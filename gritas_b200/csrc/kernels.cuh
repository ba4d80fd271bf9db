// kernels.cuh -- sm_100a device code of the GrITAS gridding hot path.
//
// classify4        box index + QC filter, shared by every accumulate kernel
//                                reference m_gritas_grids.f:405-461, gritas.f:562-587, 710-739
// k_accum_fast     direct fast accumulate (warp-aggregated fp64 RED)      m_gritas_grids.f:432-441, 465-470
// k_list_obs       K1 of the tiled fast path: one 32-byte entry per accepted observation in its tile's list
// k_tile_final     K2f / K2t / K2x / K3x: a tile's list accumulated in shared memory (counting sort, no fp64
//                  atomics), then finalized / flushed to the records / exported for or imported from the
//                  reduction across ranks                                  same lines and m_gritas_grids.f:566-654
// k_box_keys + CUB sort + k_segment_sum   deterministic accumulate: stable sort by box key, in-order segment sums
// k_finalize3/2    K3: finalize from the records                           m_gritas_grids.f:566-654, 1178-1341 (3CH)
// k_minmax         GrITAS_MinMax_                                          m_gritas_grids.f:1019-1167
// k_sortkey_*      the pre-sort of gritas.f:499-508 (with CUB radix sort passes on the host side)
//
// HBM layout of the accumulators: one record of RS doubles per box,
//   nr=1: {n, Sx, Sxx, pad}            RS=4 (32 B = one DRAM sector)
//   nr=2: {n, Sx1, Sxx1, Sx2, Sxx2, Sx1x2, pad, pad}   RS=8
//   nr=3: {n, Sx1, Sxx1, Sx2, Sxx2, Sx3, Sxx3, pad}    RS=8
// with the LEVEL index innermost: rec3 = k + km*(i + im*(j + jm*l)); surface boxes follow at
// nbox3 + i + im*(j + jm*l).  A sounding / a radiance footprint (same lat/lon, many levels or
// channels) therefore lands in consecutive records.  K3 transposes back to the reference's
// column-major (i,j,k,l,ir) arrays through shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <assert.h>

// -DGRITAS_B200_BOUNDS: every data-dependent index into shared memory, the tile lists, the records and the output planes is
// asserted (a debug build for the GPU tests; compute-sanitizer is not available on every pool).  Off by default: no code.
#ifdef GRITAS_B200_BOUNDS
#define GB_ASSERT(c) assert(c)
#else
#define GB_ASSERT(c) ((void)0)
#endif

namespace gb200 {

constexpr unsigned KEY_NONE = 0xFFFFFFFFu;
constexpr unsigned long long BAD_NONE = ~0ull;
constexpr double AMISS = 1.0e15;  // m_gritas_grids.f:72
constexpr int ACC_THREADS = 256;
constexpr int OBS_PER_THREAD = 4;

struct alignas(16) Status {
  unsigned long long first_bad;  // (call_seq << 32) | obs index of the first level miss, or BAD_NONE
  double bad_lev;
  unsigned near_pairs;  // locality probe of the direct kernel: consecutive accepted observations whose records
  unsigned pairs;       // lie within 64 of each other / all consecutive accepted pairs
};

struct GridDesc {
  int im, jm, km, nlevs, l2d, l3d, nr, rs;
  int kxmax, ktmax;
  int contiguous;  // p1(k+1)==p2(k), strictly descending: branch-free lower bound is exact
  int p2_in_smem;
  double dlon, dlat, p1_top;
  double rdlon, rdlat;   // 1/dlon, 1/dlat: fast path of the box index (exactness guarded, see grid_index)
  double rim, rjm;       // 1/im, 1/jm: column -> (i, j, l) in the tile kernel without integer divisions
  // level LUT (branch-free level lookup): cell = (high word of lev >> lut_shift) - lut_cell0 over
  // [0, lut_ncell); lut[cell] = number of interfaces at or above the cell's upper edge; at most
  // LUT_STEPS interfaces fall inside any cell.  lut_ncell == 0: binary search instead.
  int lut_shift, lut_cell0, lut_ncell;
  const unsigned short *lut;
  unsigned nbox3, nbox2;
  const int *table;
  const double *p1, *p2;
  const double *mask;    // optional (im, jm) fraction field of gritas_maskout (m_gritas_masks.F90:137-186), or null
  double *acc;
};

struct ObsCols {
  const double *lat, *lon, *lev;
  const double *d[3];        // residual source columns (del(:,ir), or omf / oma / obs / xvec / xm)
  const double *obs, *xvec, *omf;  // only for scale_res and the expression post-operations
  // general residual expressions of gritas.f:588-704 (FilterDesc::expr): del(:,ir) = sa*d[ir] [+/- db[ir]] [+ dc[ir]]
  const double *db[3], *dc[3];
  signed char sa[3], sb[3];
  const int *kx, *kt;
  const int *flag;           // ob_flag (plain entry) or qcexcl (ODS entry); may be null (all 0)
  const int *time;           // may be null
  int *flag_out;             // may be null
  int n;
  int aligned;               // every non-null column pointer is 16-byte aligned
  int flag_rw;               // flag and flag_out are the same array (GrITAS_Accum's inout ob_flag): no non-coherent loads
};

struct FilterDesc {
  int ods;        // 1: `flag` is qcexcl and the gritas.f:710-739 filter is applied
  int ioptn, t1, t2, use_time, use_lon;
  double lona, lonb;
  int x_passive, x_pre_bad;
  int scale_res;  // gritas.f:564-572
  int expr;       // residual columns are expressions (ObsCols::db/dc/sa/sb) instead of plain columns
  int post;       // 1: every column / xvec where xvec > 0 else AMISS (gritas.f:692-701); 2: column 1 / xvec (:738)
};

// Staging area of the tiled fast path.  The record grid is cut into tiles of
// 2^tile_shift consecutive records; K1 appends every accepted observation as one (record key, residuals)
// entry to the list of its tile; K2t / K2f later accumulate one tile at a time in shared memory.  Lists
// have a fixed capacity; entries that do not fit are accumulated directly.
struct StageDesc {
  unsigned *fill;        // [ntiles] tickets handed out since the last flush (may exceed cap: clamp)
  unsigned char *lists;  // [cap / LIST_G][ntiles][LIST_G] entries of ListEntry<nr>::BYTES bytes, see list_slot()
  unsigned cap, ntiles;
  int tile_shift;
  unsigned det_r0;       // deterministic tile kernel: runs longer than this are ordered by a warp's bitonic network, shorter ones by counting
  unsigned *dirty;       // [ntiles] set when an entry of the tile was accumulated directly (list overflow): the tile's
                         // records are no longer all zero
};

// Sharded multi-rank Norm: rank r finalizes the records [own_lo, own_hi) and pulls the parked
// entries of those tiles from every rank's lists over NVLink peer memory (CUDA IPC mappings of the peers' buffers) --
// the exchange step of the path fused into the tile kernel instead of a collective on the sum grids.
constexpr int MAX_PEERS = 8;
struct PeerSet {
  int n;                                  // ranks (sources); lists[self] is this rank's own buffer
  unsigned own_lo, own_hi;                // records this rank finalizes (from `bounds` when that is set)
  int self;
  const unsigned *bounds;                 // [n + 1] record boundaries of the ranks' slabs, computed on the device from the fill
                                          // counts (k_slab_bounds), or null
  const unsigned char *lists[MAX_PEERS];  // every rank's tile lists (same ntiles / cap / tile_shift on all ranks)
  const unsigned *fill[MAX_PEERS];        // local snapshot of every rank's [tickets | dirty words | records-touched word]
  const double *acc[MAX_PEERS];           // every rank's records (read only where a rank accumulated directly)
  const double *red[MAX_PEERS];           // every rank's exported tile sums (compact planes), for the planes exchange
};

// ---------------------------------------------------------------------------------------------
// Fortran NINT (round half away from zero), exact: x - trunc(x) is exact in fp64.  Out-of-range
// and NaN saturate (undefined in Fortran; pinned identically in oracle/gritas_oracle.c).
__device__ __forceinline__ long long nint_sat(double x) {
  if (!(x == x)) return -2147483648ll;
  if (x >= 2147483647.0) return 2147483647ll;
  if (x <= -2147483648.0) return -2147483648ll;
  const double t = trunc(x);
  const double f = __dsub_rn(x, t);
  long long r = (long long)t;
  r += (f >= 0.5) ? 1 : 0;
  r -= (f <= -0.5) ? 1 : 0;
  return r;
}

// 1 + NINT((x - xmin)/d) needs the correctly rounded quotient only when it lies within rounding
// distance of a half-integer.  q~ = (x - xmin) * (1/d) differs from fl((x - xmin)/d) by at most 3 ulp
// (< 1e-9 for |q| < 2^20); unless q~ is within 1e-6 of a half-integer both round to the same integer.
// Everything else (ties, huge values, NaN) takes the IEEE division: results are bit-identical to
// m_gritas_grids.f:420, 423 in all cases.
// Returned values are clipped to +-2^30: anything that far out is clamped to the grid edge by the
// caller exactly as the unclipped value would be (m_gritas_grids.f:425-426).
__device__ __noinline__ int grid_index_exact(double num, double d) {   // rare: one copy, out of line
  const long long r = nint_sat(__ddiv_rn(num, d));
  return (int)min(1073741824ll, max(-1073741824ll, r));
}

__device__ __forceinline__ int grid_index(double x, double xmin, double d, double rinv) {
  const double num = __dsub_rn(x, xmin);
  const double qa = __dmul_rn(num, rinv);
  const int k = __double2int_rn(qa);
  const double diff = __dsub_rn(qa, (double)k);
  if (fabs(diff) < 0.499999 && fabs(qa) < 1048576.0) return k;
  return grid_index_exact(num, d);
}

// gritas.f:710-739.  Returns the ob_flag the reference would hand to GrITAS_Accum_.
__device__ __forceinline__ int qc_filter(const FilterDesc &f, int qcexcl, int time, double lon) {
  const bool passive = (qcexcl == f.x_passive) | (qcexcl == f.x_pre_bad);
  const bool want = (f.ioptn > 0) & passive;
  const bool tsel = !f.use_time | ((f.t1 <= time) & (time <= f.t2));
  const bool lsel = !f.use_lon | ((f.lona <= lon) & (lon <= f.lonb));
  return (tsel & lsel & ((qcexcl == 0) | want)) ? 0 : 1;
}

// m_gritas_grids.f:450-456.  0-based level, or -1 ("could not find level").
constexpr int LUT_STEPS = 2;

// Host: builds the level LUT for a contiguous table (p1, p2 from GrITAS_Pint): cells of 2^-m octave on
// the high word of lev, with the smallest m in [3, 10] for which no cell holds more than LUT_STEPS
// interfaces.  Returns the number of cells (0: no LUT, use the binary search).
inline int build_level_lut(int nlevs, const double *p1, const double *p2, unsigned short *out, int out_cap, int *shift_out,
                           int *cell0_out) {
  if (nlevs < 2 || nlevs > 60000 || !(p2[nlevs - 2] > 0.0)) return 0;
  const int len = nlevs - 1;
  auto hiword = [](double x) { unsigned long long u; memcpy(&u, &x, 8); return (unsigned)(u >> 32); };
  auto fromhi = [](unsigned hw) { unsigned long long u = (unsigned long long)hw << 32; double x; memcpy(&x, &u, 8); return x; };
  const double vlo = p2[len - 1], vhi = p1[0] > p2[0] ? p1[0] : p2[0];
  for (int mbits = 3; mbits <= 10; ++mbits) {
    const int shift = 20 - mbits;
    const unsigned c0 = hiword(vlo) >> shift, c1 = hiword(vhi) >> shift;
    if (c1 < c0 || (int)(c1 - c0 + 1) > out_cap) continue;
    const int ncell = (int)(c1 - c0 + 1);
    int kk = 0;  // interfaces are descending: walk the cells from the top
    bool ok = true;
    for (int cell = ncell - 1; cell >= 0 && ok; --cell) {
      const double ub = fromhi((c0 + (unsigned)cell + 1u) << shift), lb = fromhi((c0 + (unsigned)cell) << shift);
      while (kk < len && p2[kk] >= ub) ++kk;        // kk = number of interfaces >= ub
      out[cell] = (unsigned short)kk;
      int inside = 0;
      for (int q = kk; q < len && p2[q] >= lb; ++q) ++inside;
      if (inside > LUT_STEPS) ok = false;
    }
    if (ok) {
      *shift_out = shift;
      *cell0_out = (int)c0;
      return ncell;
    }
  }
  return 0;
}

__device__ __forceinline__ int find_level(const GridDesc &g, const double *__restrict__ p2s,
                                          const unsigned short *__restrict__ luts, double lev) {
  const int nl = g.nlevs;
  if (g.contiguous) {
    // level = #{kk < nl-1 : lev <= p2(kk)}; p2 strictly descending so the predicate holds on a prefix.
    const int len = nl - 1;
    int pos = 0;
    if (g.lut_ncell > 0) {
      // branch-free: table on the exponent + leading mantissa bits of lev, then LUT_STEPS exact compares
      const unsigned hi = (unsigned)__double2hiint(lev);
      int cell = (int)(hi >> g.lut_shift) - g.lut_cell0;   // negative lev / NaN / inf land beyond the table
      cell = min(max(cell, 0), g.lut_ncell - 1);
      pos = (int)luts[cell];
#pragma unroll
      for (int s = 0; s < LUT_STEPS; ++s) pos += (pos < len && lev <= p2s[pos]) ? 1 : 0;
    } else {
      int step = 1;
      while ((step << 1) <= len) step <<= 1;
      for (; step > 0; step >>= 1) {
        const int np = pos + step;
        if (np <= len && lev <= p2s[np - 1]) pos = np;
      }
    }
    const double top = (pos == 0) ? g.p1_top : p2s[pos - 1];
    return (lev <= top && lev > p2s[pos]) ? pos : -1;
  }
  for (int kk = 0; kk < nl; ++kk)
    if (lev <= __ldg(g.p1 + kk) && lev > __ldg(g.p2 + kk)) return kk;
  return -1;
}

// m_gritas_grids.f:409-426 + level search.  outcome: 0 accumulate, 1 not in table, 2 level miss.
__device__ __forceinline__ unsigned box_key(const GridDesc &g, const double *__restrict__ p2s,
                                            const unsigned short *__restrict__ luts, double lat, double lon,
                                            double lev, int kx, int kt, int &outcome) {
  int l = 0;
  if (kx >= 1 && kx <= g.kxmax && kt >= 1 && kt <= g.ktmax)
    l = __ldg(g.table + (size_t)(kx - 1) + (size_t)g.kxmax * (size_t)(kt - 1));
  const bool surface = l < 0;
  l = abs(l);
  if (l <= 0) { outcome = 1; return KEY_NONE; }
  int ii = 1 + grid_index(lon, -180.0, g.dlon, g.rdlon);
  if (ii == g.im + 1) ii = 1;
  if (ii == 0) ii = g.im;
  int jj = 1 + grid_index(lat, -90.0, g.dlat, g.rdlat);
  ii = min(g.im, max(1, ii));
  jj = min(g.jm, max(1, jj));
  const unsigned i = (unsigned)(ii - 1), j = (unsigned)(jj - 1);
  outcome = 0;
  if (surface) return g.nbox3 + i + (unsigned)g.im * (j + (unsigned)g.jm * (unsigned)(l - 1));
  const int k = find_level(g, p2s, luts, lev);
  if (k < 0) { outcome = 2; return KEY_NONE; }
  return (unsigned)k + (unsigned)g.km * (i + (unsigned)g.im * (j + (unsigned)g.jm * (unsigned)(l - 1)));
}

// gritas.f:564-572 residual scaling (applied to every selected column, as the reference does for
// column 1; column 2 of the omf+oma mode is the second reference run).
__device__ __forceinline__ double scale_residual(const FilterDesc &f, const ObsCols &c, double d, int n) {
  if (f.scale_res == 1) return __ddiv_rn(d, __dsqrt_rn(__ldg(c.xvec + n)));
  if (f.scale_res == 2) return __ddiv_rn(d, __ldg(c.obs + n));
  if (f.scale_res == 3) return __ddiv_rn(d, __dsub_rn(__ldg(c.obs + n), __ldg(c.omf + n)));
  return d;
}

// Residual column ir of observation n from its primary source value d: the expressions of gritas.f:588-704
// evaluated left to right exactly as written there, then the optional scaling.
__device__ __forceinline__ double residual_value(const FilterDesc &f, const ObsCols &c, int ir, double d, int n) {
  if (f.expr) {
    if (c.sa[ir] < 0) d = -d;
    if (c.db[ir]) {
      const double b = __ldg(c.db[ir] + n);
      d = c.sb[ir] > 0 ? __dadd_rn(d, b) : __dsub_rn(d, b);
    }
    if (c.dc[ir]) d = __dadd_rn(d, __ldg(c.dc[ir] + n));
    if (f.post == 1) {
      const double x = __ldg(c.xvec + n);
      d = x > 0.0 ? __ddiv_rn(d, x) : AMISS;
    }
    if (f.post == 2 && ir == 0) d = __ddiv_rn(d, __ldg(c.xvec + n));
  }
  return f.scale_res ? scale_residual(f, c, d, n) : d;
}

// Level miss (m_gritas_grids.f:458-459): keep the EARLIEST (call, observation) together with its level.
// The pair {first_bad, bad_lev} is the first 16 bytes of Status and is replaced with one 128-bit CAS, so
// the level always belongs to the recorded observation -- no per-CTA fence / last-block pass needed.
__device__ __noinline__ void report_miss(Status *st, unsigned call_seq, int n, double lev) {
  const unsigned long long mine = ((unsigned long long)call_seq << 32) | (unsigned)n;
  unsigned long long cur_id = *reinterpret_cast<volatile unsigned long long *>(&st->first_bad);
  unsigned long long cur_lev = *reinterpret_cast<volatile unsigned long long *>(&st->bad_lev);
  const unsigned long long my_lev = (unsigned long long)__double_as_longlong(lev);
  while (mine < cur_id) {
    unsigned long long got_id, got_lev;
    asm volatile(
        "{\n\t.reg .b128 cmp, val, old;\n\t"
        "mov.b128 cmp, {%2, %3};\n\t"
        "mov.b128 val, {%4, %5};\n\t"
        "atom.global.cas.b128 old, [%6], cmp, val;\n\t"
        "mov.b128 {%0, %1}, old;\n\t}"
        : "=l"(got_id), "=l"(got_lev)
        : "l"(cur_id), "l"(cur_lev), "l"(mine), "l"(my_lev), "l"(st)
        : "memory");
    if (got_id == cur_id && got_lev == cur_lev) break;
    cur_id = got_id;
    cur_lev = got_lev;
  }
}

// L2 residency policy: the observation columns are read exactly once (evict first), so that they do not push the
// open tails of the tile lists out of L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// streaming 128-bit loads that do not pollute L1 and leave L2 first
__device__ __forceinline__ double2 ldg_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
               : "=d"(r.x), "=d"(r.y)
               : "l"(p), "l"(l2_policy_evict_first()));
  return r;
}
__device__ __forceinline__ int4 ldg_stream4(const int *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(l2_policy_evict_first()));
  return r;
}
// same for a column the kernel also writes (ob_flag inout): ld.global.nc would be undefined there
__device__ __forceinline__ int4 ldg_rw4(const int *p) {
  int4 r;
  asm volatile("ld.global.cs.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ int ld_rw(const int *p) {
  int r;
  asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
  return r;
}
template <int NR>
struct Obs4 {
  double lat[4], lon[4], lev[4], d[NR][4];
  int kx[4], kt[4], fl[4], tm[4];
};

// Coordinates, kx/kt, flag and time of 4 consecutive observations (36-40 B per observation).
template <int NR>
__device__ __forceinline__ void load_geo4_slow(const ObsCols &c, const FilterDesc &f, int base, Obs4<NR> &o) {
  const int n = c.n;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = base + j;
    const bool in = i < n;
    const int ii = in ? i : 0;
    o.lat[j] = in ? __ldg(c.lat + ii) : 0.0;
    o.lon[j] = in ? __ldg(c.lon + ii) : 0.0;
    o.lev[j] = in ? __ldg(c.lev + ii) : 0.0;
    o.kx[j] = in ? __ldg(c.kx + ii) : 0;
    o.kt[j] = in ? __ldg(c.kt + ii) : 0;
    o.fl[j] = (in && c.flag) ? (c.flag_rw ? ld_rw(c.flag + ii) : __ldg(c.flag + ii)) : 0;
    o.tm[j] = (in && f.use_time) ? __ldg(c.time + ii) : 0;
  }
}

template <int NR>
__device__ __forceinline__ void load_geo4(const ObsCols &c, const FilterDesc &f, int base, Obs4<NR> &o) {
  const int n = c.n;
  if (c.aligned && base + 3 < n) {
    double2 a, b;
    a = ldg_stream2(c.lat + base); b = ldg_stream2(c.lat + base + 2);
    o.lat[0] = a.x; o.lat[1] = a.y; o.lat[2] = b.x; o.lat[3] = b.y;
    a = ldg_stream2(c.lon + base); b = ldg_stream2(c.lon + base + 2);
    o.lon[0] = a.x; o.lon[1] = a.y; o.lon[2] = b.x; o.lon[3] = b.y;
    a = ldg_stream2(c.lev + base); b = ldg_stream2(c.lev + base + 2);
    o.lev[0] = a.x; o.lev[1] = a.y; o.lev[2] = b.x; o.lev[3] = b.y;
    int4 v = ldg_stream4(c.kx + base);
    o.kx[0] = v.x; o.kx[1] = v.y; o.kx[2] = v.z; o.kx[3] = v.w;
    v = ldg_stream4(c.kt + base);
    o.kt[0] = v.x; o.kt[1] = v.y; o.kt[2] = v.z; o.kt[3] = v.w;
    if (c.flag) {
      v = c.flag_rw ? ldg_rw4(c.flag + base) : ldg_stream4(c.flag + base);
      o.fl[0] = v.x; o.fl[1] = v.y; o.fl[2] = v.z; o.fl[3] = v.w;
    } else {
      o.fl[0] = o.fl[1] = o.fl[2] = o.fl[3] = 0;
    }
    if (f.use_time) {
      v = ldg_stream4(c.time + base);
      o.tm[0] = v.x; o.tm[1] = v.y; o.tm[2] = v.z; o.tm[3] = v.w;
    } else {
      o.tm[0] = o.tm[1] = o.tm[2] = o.tm[3] = 0;
    }
  } else {
    load_geo4_slow<NR>(c, f, base, o);   // ragged tail / unaligned caller arrays
  }
}

// Residual columns of 4 consecutive observations (8 B per observation and column).
template <int NR>
__device__ __forceinline__ void load_del4_slow(const ObsCols &c, int base, double (&d)[NR][4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = base + j;
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) d[ir][j] = (i < c.n) ? __ldg(c.d[ir] + i) : 0.0;
  }
}

template <int NR>
__device__ __forceinline__ void load_del4(const ObsCols &c, int base, double (&d)[NR][4]) {
  if (c.aligned && base + 3 < c.n) {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      const double2 a = ldg_stream2(c.d[ir] + base), b = ldg_stream2(c.d[ir] + base + 2);
      d[ir][0] = a.x; d[ir][1] = a.y; d[ir][2] = b.x; d[ir][3] = b.y;
    }
  } else {
    load_del4_slow<NR>(c, base, d);
  }
}

template <int NR>
__device__ __forceinline__ void load_obs4(const ObsCols &c, const FilterDesc &f, int base, Obs4<NR> &o) {
  load_geo4<NR>(c, f, base, o);
  load_del4<NR>(c, base, o.d);
}

// Per-observation front half shared by the fast kernel and the key kernel: filter, table lookup,
// box index.  Returns the record key (KEY_NONE if nothing is to be accumulated) and the ob_flag
// the reference would leave behind.
struct LevTab {   // level tables as the kernels see them (shared memory when they fit)
  const double *p2;
  const unsigned short *lut;
};

__device__ __forceinline__ unsigned classify(const GridDesc &g, const FilterDesc &f, const LevTab p2s,
                                             double lat, double lon, double lev, int kx, int kt, int flag_in,
                                             int tm, int n, Status *st, unsigned call_seq, int &flag_after) {
  int fl = flag_in;
  if (f.ods) fl = qc_filter(f, flag_in, tm, lon);
  if (g.mask && fl == 0) {  // gritas_maskout (gritas.f:745): same box formula, fraction < 0.99 -> flagged
    int ii = 1 + grid_index(lon, -180.0, g.dlon, g.rdlon);
    if (ii == g.im + 1) ii = 1;
    if (ii == 0) ii = g.im;
    ii = min(g.im, max(1, ii));
    const int jj = min(g.jm, max(1, 1 + grid_index(lat, -90.0, g.dlat, g.rdlat)));
    if (__ldg(g.mask + (size_t)(ii - 1) + (size_t)g.im * (size_t)(jj - 1)) < 0.99) fl = 1;
  }
  flag_after = fl;
  if (fl != 0) return KEY_NONE;  // m_gritas_grids.f:406-407
  int outcome;
  const unsigned key = box_key(g, p2s.p2, p2s.lut, lat, lon, lev, kx, kt, outcome);
  if (outcome == 1) flag_after = 1;  // m_gritas_grids.f:413-416
  if (outcome == 2) report_miss(st, call_seq, n, lev);
  return key;
}

// classify() for 4 consecutive observations, written step by step over the four so that the
// independent chains (table lookup, two box indices, level lookup) of different observations overlap.
// The only control flow is one rarely taken branch for the exact-division cases and one for level
// misses.  FMODE: 0 = flags as given (GrITAS_Accum entry), 1 = QC filter without time / longitude
// window (the usual ODS case), 2 = full filter of gritas.f:710-739.  Same results as classify().
template <int NR, int FMODE>
__device__ __forceinline__ void classify4(const GridDesc &g, const FilterDesc &f, const LevTab t, const Obs4<NR> &o,
                                          int base, int n, Status *st, unsigned call_seq, unsigned (&key)[4],
                                          int (&flag_after)[4]) {
  int fl[4], l[4], ii[4], jj[4], k[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (FMODE == 0) {
      fl[j] = o.fl[j];
    } else if (FMODE == 1) {
      const bool passive = (o.fl[j] == f.x_passive) | (o.fl[j] == f.x_pre_bad);
      fl[j] = ((o.fl[j] == 0) | ((f.ioptn > 0) & passive)) ? 0 : 1;
    } else {
      fl[j] = qc_filter(f, o.fl[j], o.tm[j], o.lon[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned kx1 = (unsigned)(o.kx[j] - 1), kt1 = (unsigned)(o.kt[j] - 1);
    const bool inr = kx1 < (unsigned)g.kxmax && kt1 < (unsigned)g.ktmax;
    const unsigned idx = inr ? kx1 + (unsigned)g.kxmax * kt1 : 0u;
    const int v = __ldg(g.table + idx);
    l[j] = inr ? v : 0;                            // out-of-range kx/kt: "not in table" (m_gritas_grids.f:409)
  }
  // box indices: reciprocal multiply for all eight coordinates, exactness checked once (see grid_index)
  double num[8];
  unsigned inexact = 0u;   // bit q: coordinate q needs the IEEE division
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    num[j] = __dsub_rn(o.lon[j], -180.0);
    num[4 + j] = __dsub_rn(o.lat[j], -90.0);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const double qa = __dmul_rn(num[q], q < 4 ? g.rdlon : g.rdlat);
    const int kq = __double2int_rn(qa);
    const double diff = __dsub_rn(qa, (double)kq);
    inexact |= (fabs(diff) < 0.499999 && fabs(qa) < 1048576.0) ? 0u : (1u << q);
    if (q < 4) ii[q] = kq; else jj[q - 4] = kq;
  }
  if (inexact) {   // ties, huge values, NaN: the IEEE division decides (rare)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (inexact & (1u << q)) {
        const int kq = grid_index_exact(num[q], q < 4 ? g.dlon : g.dlat);
        if (q < 4) ii[q] = kq; else jj[q - 4] = kq;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = find_level(g, t.p2, t.lut, o.lev[j]);                 // :450-456
  bool any_miss = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int i2 = ii[j] + 1;                                                                      // :420
    i2 = (i2 == g.im + 1) ? 1 : i2;                                                          // :421
    i2 = (i2 == 0) ? g.im : i2;                                                              // :422
    i2 = min(g.im, max(1, i2));                                                              // :425
    const int j2 = min(g.jm, max(1, jj[j] + 1));                                             // :423, 426
    if (g.mask && fl[j] == 0 && base + j < n)   // gritas_maskout (gritas.f:745, m_gritas_masks.F90:164-183)
      fl[j] = (__ldg(g.mask + (size_t)(i2 - 1) + (size_t)g.im * (size_t)(j2 - 1)) < 0.99) ? 1 : 0;
    const bool live = (base + j < n) && fl[j] == 0;                                         // :406-407
    const bool surface = l[j] < 0;                                                           // :410
    const int la = abs(l[j]);
    const bool intab = la > 0;                                                               // :413
    const unsigned cell = (unsigned)(i2 - 1) + (unsigned)g.im * ((unsigned)(j2 - 1) + (unsigned)g.jm * (unsigned)(la - 1));
    const unsigned key2 = g.nbox3 + cell;
    const unsigned key3 = (unsigned)k[j] + (unsigned)g.km * cell;
    const bool ok = live && intab && (surface || k[j] >= 0);
    key[j] = ok ? (surface ? key2 : key3) : KEY_NONE;
    flag_after[j] = (base + j < n) ? ((live && !intab) ? 1 : fl[j]) : 0;                     // :413-416
    any_miss |= live && intab && !surface && k[j] < 0;
  }
  if (any_miss) {                                                                            // :458-459
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool live = (base + j < n) && fl[j] == 0;
      if (live && l[j] > 0 && k[j] < 0) report_miss(st, call_seq, base + j, o.lev[j]);
    }
  }
}

template <int NR>
__device__ __forceinline__ void red_record(double *__restrict__ rec, double cnt, const double (&sx)[NR],
                                           const double (&sq)[NR], double sxy) {
  atomicAdd(rec, cnt);
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) {
    atomicAdd(rec + 1 + 2 * ir, sx[ir]);
    atomicAdd(rec + 2 + 2 * ir, sq[ir]);
  }
  if (NR == 2) atomicAdd(rec + 5, sxy);
}

// Level tables into shared memory (p2 doubles, then the LUT); includes the barrier.
__host__ __device__ inline size_t levtab_smem_bytes(const GridDesc &g) {
  if (!g.p2_in_smem) return 0;
  return sizeof(double) * (size_t)g.nlevs + sizeof(unsigned short) * (size_t)((g.lut_ncell + 3) & ~3);
}

__device__ __forceinline__ LevTab load_p2(const GridDesc &g, double *s_p2) {
  LevTab t;
  t.p2 = g.p2;
  t.lut = g.lut;
  if (g.p2_in_smem) {
    unsigned short *s_lut = reinterpret_cast<unsigned short *>(s_p2 + g.nlevs);
    for (int k = threadIdx.x; k < g.nlevs; k += blockDim.x) s_p2[k] = g.p2[k];
    for (int k = threadIdx.x; k < g.lut_ncell; k += blockDim.x) s_lut[k] = g.lut[k];
    t.p2 = s_p2;
    t.lut = s_lut;
  }
  __syncthreads();
  return t;
}

// ---------------------------------------------------------------------------------------------
// K1+K2a fused: stream the columns, classify, accumulate with fp64 RED into the record grid.
// AGG: lanes of a warp that hit the same box (__match_any_sync) are summed in registers first and
// only the lowest lane issues the atomics.
template <int NR, bool AGG, int FMODE, bool PROBE = false>
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_fast(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq) {
  extern __shared__ double s_p2[];
  const LevTab p2s = load_p2(g, s_p2);
  const int nvec = (c.n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
  const int nvec_w = (nvec + 31) & ~31;  // whole warps iterate together (match/shfl need all lanes)
  const int lane = threadIdx.x & 31;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec_w; v += gridDim.x * blockDim.x) {
    const int base = v * OBS_PER_THREAD;
    Obs4<NR> o;
    int fa[4] = {0, 0, 0, 0};
    unsigned k4[4] = {KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE};
    if (v < nvec) {
      load_obs4<NR>(c, f, base, o);
      classify4<NR, FMODE>(g, f, p2s, o, base, c.n, st, call_seq, k4, fa);
    }
    if (PROBE) {
      // How local is this data?  Soundings / radiance footprints put consecutive observations into neighbouring
      // records (the direct RED path coalesces); scattered data does not (the tiled path wins).
      unsigned nearc = 0, pairc = 0;
#pragma unroll
      for (int j = 1; j < 4; ++j) {
        const bool both = k4[j] != KEY_NONE && k4[j - 1] != KEY_NONE;
        const unsigned dist = k4[j] > k4[j - 1] ? k4[j] - k4[j - 1] : k4[j - 1] - k4[j];
        pairc += both ? 1u : 0u;
        nearc += (both && dist < 64u) ? 1u : 0u;
      }
      nearc = __reduce_add_sync(0xffffffffu, nearc);
      pairc = __reduce_add_sync(0xffffffffu, pairc);
      if (lane == 0 && pairc) { atomicAdd(&st->near_pairs, nearc); atomicAdd(&st->pairs, pairc); }
    }
    // the probe only looks: the call is then accumulated on the path the host picks (nothing reaches the records, so a
    // tiled month's Norm does not have to read them)
    if constexpr (!PROBE) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned key = k4[j];
      double sx[NR], sq[NR], sxy = 0.0;
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) {
        double d = o.d[ir][j];
        if ((f.scale_res | f.expr) && key != KEY_NONE) d = residual_value(f, c, ir, d, base + j);
        sx[ir] = d;
        sq[ir] = __dmul_rn(d, d);
      }
      if (NR == 2) sxy = __dmul_rn(sx[0], sx[1]);
      double cnt = 1.0;
      bool issue = key != KEY_NONE;
      if (AGG) {
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const bool multi = issue && (peers & (peers - 1));
        if (__any_sync(0xffffffffu, multi)) {
          const int leader = __ffs(peers) - 1;
          const int npeer = issue ? __popc(peers) : 1;
          const int maxc = __reduce_max_sync(0xffffffffu, npeer);
          unsigned others = peers & ~(1u << leader);
          for (int it = 1; it < maxc; ++it) {
            const int src = others ? (__ffs(others) - 1) : lane;
            others &= others - 1;
            const bool take = (lane == leader) && (it < npeer);
#pragma unroll
            for (int ir = 0; ir < NR; ++ir) {
              const double a = __shfl_sync(0xffffffffu, sx[ir], src);
              const double b = __shfl_sync(0xffffffffu, sq[ir], src);
              if (take) { sx[ir] = __dadd_rn(sx[ir], a); sq[ir] = __dadd_rn(sq[ir], b); }
            }
            if (NR == 2) {
              const double a = __shfl_sync(0xffffffffu, sxy, src);
              if (take) sxy = __dadd_rn(sxy, a);
            }
          }
          cnt = (double)npeer;
          issue = issue && (lane == leader);
        }
      }
      if (issue) red_record<NR>(g.acc + (size_t)key * (size_t)g.rs, cnt, sx, sq, sxy);
    }
    if (c.flag_out && v < nvec) {
      if (c.aligned && base + 3 < c.n) {
        *reinterpret_cast<int4 *>(c.flag_out + base) = make_int4(fa[0], fa[1], fa[2], fa[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (base + j < c.n) c.flag_out[base + j] = fa[j];
      }
    }
    }   // !PROBE
  }
}

// ---------------------------------------------------------------------------------------------
// GrITAS_MinMax_ (m_gritas_grids.f:1019-1167): same filter / box index, keeps the smallest and largest
// residual of every box and the count.  Records are {n, min, max, pad}; min / max are held as order-preserving
// int64 keys so that they can be updated with native 64-bit atomicMin / atomicMax (order-free: exact and
// reproducible in every mode).  NaN residuals count but never win a comparison, as in the reference.
__device__ __forceinline__ long long f64_order_key(double x) {
  const long long b = __double_as_longlong(x);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffll);
}
__device__ __forceinline__ double f64_from_order_key(long long k) {
  return __longlong_as_double(k >= 0 ? k : (k ^ 0x7fffffffffffffffll));
}

__global__ void __launch_bounds__(256)
k_minmax_init(double *__restrict__ acc, size_t nrec) {
  for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrec; r += (size_t)gridDim.x * blockDim.x) {
    double *rec = acc + r * 4;
    rec[0] = 0.0;
    reinterpret_cast<long long *>(rec)[1] = f64_order_key(AMISS);    // gritas.f:364-366: min starts at AMISS,
    reinterpret_cast<long long *>(rec)[2] = f64_order_key(-AMISS);   // max at -AMISS
    rec[3] = 0.0;
  }
}

template <int FMODE>
__global__ void __launch_bounds__(ACC_THREADS)
k_minmax(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq) {
  extern __shared__ double s_p2[];
  const LevTab p2s = load_p2(g, s_p2);
  const int nvec = (c.n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    const int base = v * OBS_PER_THREAD;
    Obs4<1> o;
    int fa[4];
    unsigned k4[4];
    load_obs4<1>(c, f, base, o);
    classify4<1, FMODE>(g, f, p2s, o, base, c.n, st, call_seq, k4, fa);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k4[j] == KEY_NONE) continue;
      double d = o.d[0][j];
      if (f.scale_res | f.expr) d = residual_value(f, c, 0, d, base + j);
      double *rec = g.acc + (size_t)k4[j] * 4;
      atomicAdd(rec, 1.0);                                                // :1113, 1148
      if (d == d) {
        atomicMin(reinterpret_cast<long long *>(rec) + 1, f64_order_key(d));   // :1109, 1144
        atomicMax(reinterpret_cast<long long *>(rec) + 2, f64_order_key(d));   // :1110, 1145
      }
    }
    if (c.flag_out) {
      if (c.aligned && base + 3 < c.n) {
        *reinterpret_cast<int4 *>(c.flag_out + base) = make_int4(fa[0], fa[1], fa[2], fa[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (base + j < c.n) c.flag_out[base + j] = fa[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1 alone (deterministic path): keys + identity permutation for the stable sort.
__global__ void __launch_bounds__(ACC_THREADS)
k_box_keys(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq,
           unsigned *__restrict__ keys, unsigned *__restrict__ idx, unsigned nrec) {
  extern __shared__ double s_p2[];
  const LevTab p2s = load_p2(g, s_p2);
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < c.n; n += gridDim.x * blockDim.x) {
    const int fl = c.flag ? (c.flag_rw ? ld_rw(c.flag + n) : __ldg(c.flag + n)) : 0;
    const int tm = f.use_time ? __ldg(c.time + n) : 0;
    int fa;
    const unsigned key = classify(g, f, p2s, __ldg(c.lat + n), __ldg(c.lon + n), __ldg(c.lev + n), __ldg(c.kx + n),
                                  __ldg(c.kt + n), fl, tm, n, st, call_seq, fa);
    keys[n] = key < nrec ? key : nrec;  // rejected observations sort last within key_bits
    idx[n] = (unsigned)n;
    if (c.flag_out) c.flag_out[n] = fa;
  }
}

// K2b: one thread per run of equal keys in the stably sorted list.  The run is added to the
// record one observation at a time, in input order, continuing from the value already there:
// the same sequence of fp64 additions as the reference loop m_gritas_grids.f:405-474.
template <int NR>
__global__ void __launch_bounds__(ACC_THREADS)
k_segment_sum(const GridDesc g, const ObsCols c, const FilterDesc f, const unsigned *__restrict__ keys,
              const unsigned *__restrict__ idx, int n, unsigned nrec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const unsigned key = keys[t];
  if (key >= nrec) return;
  if (t > 0 && keys[t - 1] == key) return;  // not a segment head
  double *rec = g.acc + (size_t)key * (size_t)g.rs;
  double cnt = rec[0], sx[NR], sq[NR], sxy = (NR == 2) ? rec[5] : 0.0;
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) { sx[ir] = rec[1 + 2 * ir]; sq[ir] = rec[2 + 2 * ir]; }
  for (int u = t; u < n && keys[u] == key; ++u) {
    const int i = (int)idx[u];
    double d[NR];
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      d[ir] = residual_value(f, c, ir, __ldg(c.d[ir] + i), i);
      sx[ir] = __dadd_rn(sx[ir], d[ir]);
      sq[ir] = __dadd_rn(sq[ir], __dmul_rn(d[ir], d[ir]));
    }
    if (NR == 2) sxy = __dadd_rn(sxy, __dmul_rn(d[0], d[1]));
    cnt = __dadd_rn(cnt, 1.0);
  }
  rec[0] = cnt;
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) { rec[1 + 2 * ir] = sx[ir]; rec[2 + 2 * ir] = sq[ir]; }
  if (NR == 2) rec[5] = sxy;
}

// ---------------------------------------------------------------------------------------------
// Tiled fast path, part 1 (K1): every accepted observation becomes ONE entry in the list of its tile.
//
// An entry is one aligned 32-byte sector {key u32, -, d1 f64, d2 f64 [, d3]} (16 bytes {key, -, d1} at nr = 1) and
// is written with a single 256-bit store at the position a ticket (atomicAdd on the tile's fill counter) hands out.
// Measured on B200 (profiles/r1_microbench_one_level.txt): scattered stores run at ~40 G requests/s whether they
// carry 4 or 32 bytes, tickets on 13 K counters at ~65 G/s, and both behind the 52 B/obs column stream give
// 47 G obs/s -- faster than grouping the entries in shared memory first (two partition kernels per file, an
// L2-resident scratch and ~13 % padding, 31 G obs/s), because that path was bound by instruction issue, not by
// memory.  Lanes of a warp that hit the same tile share one ticket (__match_any_sync): clustered data
// (radiosonde sites, swaths) would otherwise serialise on the tile's counter.
template <int NR>
struct ListEntry {
  static constexpr int BYTES = (NR == 1) ? 16 : 32;
};

// Where entry `pos` of tile `tile` lives.  The lists are interleaved in granules of LIST_G entries (256 bytes at
// nr = 2): granule c of every tile sits next to granule c of its neighbours, so the open tails of all lists form one
// contiguous frontier that moves through memory as the month is read.  With list-major storage every tail is a
// separate DRAM row and the entry stores run at the random-sector rate (~40 G/s); interleaved they reach the
// ticket rate (~60 G/s), and a tile's list still reads back at full bandwidth (profiles/r1_microbench_one_level.txt).
constexpr unsigned LIST_G = 8;
__device__ __forceinline__ size_t list_slot(const StageDesc &sg, unsigned tile, unsigned pos) {
  return ((size_t)(pos / LIST_G) * sg.ntiles + tile) * LIST_G + (pos % LIST_G);
}

// Deterministic path: the first 8 bytes of an entry carry the box number within the tile (16 bits) and the observation's
// sequence number (48 bits: call order << 28 | index within the call) -- the tile kernel sums every box in that order.
constexpr int SEQ_IDX_BITS = 28;
template <int NR>
__device__ __forceinline__ void entry_store(unsigned char *lists, size_t idx, unsigned long long word0, const double (&d)[NR]) {
  unsigned char *p = lists + idx * (size_t)ListEntry<NR>::BYTES;
  const double k = __longlong_as_double((long long)word0);
  if (NR == 1) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(k), "d"(d[0]) : "memory");
  } else {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(k), "d"(d[0]), "d"(d[1]), "d"(d[NR - 1]) : "memory");
  }
}

// Streaming (read-once) load of entry idx; returns the key.
template <int NR>
__device__ __forceinline__ unsigned entry_load(const unsigned char *lists, size_t idx, double (&d)[NR]) {
  const unsigned char *p = lists + idx * (size_t)ListEntry<NR>::BYTES;
  double k, a, b, c;
  if (NR == 1) {
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(k), "=d"(a) : "l"(p));
    d[0] = a;
  } else {
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(k), "=d"(a), "=d"(b), "=d"(c) : "l"(p));
    d[0] = a;
    d[1] = b;
    if (NR == 3) d[NR - 1] = c;
  }
  return (unsigned)__double_as_longlong(k);
}

// Deterministic path: the entry's first word (see entry_store) and its residuals.
template <int NR>
__device__ __forceinline__ unsigned long long entry_load64(const unsigned char *lists, size_t idx, double (&d)[NR]) {
  const unsigned char *p = lists + idx * (size_t)ListEntry<NR>::BYTES;
  double k, a, b, c;
  if (NR == 1) {
    asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(k), "=d"(a) : "l"(p));
    d[0] = a;
  } else {
    asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(k), "=d"(a), "=d"(b), "=d"(c) : "l"(p));
    d[0] = a;
    d[1] = b;
    if (NR == 3) d[NR - 1] = c;
  }
  return (unsigned long long)__double_as_longlong(k);
}
__device__ __forceinline__ unsigned long long entry_word0(const unsigned char *lists, size_t idx, int bytes) {
  unsigned long long w;
  asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(w) : "l"(lists + idx * (size_t)bytes));
  return w;
}

// Direct accumulate of one entry (a full list): the same fp64 RED the direct path uses.
template <int NR>
__device__ __forceinline__ void red_entry(const GridDesc &g, unsigned key, const double (&d)[NR]) {
  double sq[NR];
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) sq[ir] = __dmul_rn(d[ir], d[ir]);
  const double sxy = (NR == 2) ? __dmul_rn(d[0], d[NR - 1]) : 0.0;
  red_record<NR>(g.acc + (size_t)key * (size_t)g.rs, 1.0, d, sq, sxy);
}

constexpr int K1_THREADS = 256;
#ifndef GB_K1_MINB
#define GB_K1_MINB 4
#endif

template <int NR, int FMODE, bool DET = false>
__global__ void __launch_bounds__(K1_THREADS, GB_K1_MINB)
k_list_obs(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq,
           const StageDesc sg, unsigned long long seq_base = 0ull) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const LevTab p2s = load_p2(g, reinterpret_cast<double *>(smem_raw));
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = (1u << lane) - 1u;
  const int nvec = (c.n + 3) >> 2;
  // whole warps iterate together (the ticket aggregation is warp-wide)
  const int nround = (nvec + K1_THREADS - 1) / K1_THREADS * K1_THREADS;
  for (int vec = blockIdx.x * K1_THREADS + threadIdx.x; vec < nround; vec += gridDim.x * K1_THREADS) {
    const int base = vec * 4;
    unsigned k4[4] = {KEY_NONE, KEY_NONE, KEY_NONE, KEY_NONE};
    if (base < c.n) {
      Obs4<NR> o;
      int fa[4];
      load_geo4<NR>(c, f, base, o);
      classify4<NR, FMODE>(g, f, p2s, o, base, c.n, st, call_seq, k4, fa);
      if (c.flag_out) {
        if (c.aligned && base + 3 < c.n) {
          *reinterpret_cast<int4 *>(c.flag_out + base) = make_int4(fa[0], fa[1], fa[2], fa[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (base + j < c.n) c.flag_out[base + j] = fa[j];
        }
      }
    }
    const bool any = (k4[0] & k4[1] & k4[2] & k4[3]) != KEY_NONE;
    if (!__any_sync(0xffffffffu, any)) continue;
    double dd[NR][4];
    if (any) load_del4<NR>(c, base, dd);
    // tickets first (four independent atomics in flight), then the stores
    unsigned pos[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned tile = (k4[j] == KEY_NONE) ? 0xFFFFFFFFu : (k4[j] >> sg.tile_shift);
      const unsigned peers = __match_any_sync(0xffffffffu, tile);
      const int leader = __ffs(peers) - 1;
      unsigned p0 = 0u;
      if ((int)lane == leader && tile != 0xFFFFFFFFu) p0 = atomicAdd(sg.fill + tile, (unsigned)__popc(peers));
      pos[j] = __shfl_sync(0xffffffffu, p0, leader) + (unsigned)__popc(peers & lt);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned key = k4[j];
      if (key == KEY_NONE) continue;
      double d[NR];
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) d[ir] = (f.scale_res | f.expr) ? residual_value(f, c, ir, dd[ir][j], base + j) : dd[ir][j];
      const unsigned tile = key >> sg.tile_shift;
      GB_ASSERT(tile < sg.ntiles && (size_t)key < (size_t)g.nbox3 + (size_t)g.nbox2);
      if (pos[j] < sg.cap) {
        GB_ASSERT(list_slot(sg, tile, pos[j]) < (size_t)sg.cap * sg.ntiles);
        const unsigned long long w0 = DET ? (((seq_base + (unsigned long long)(base + j)) << 16) | (unsigned long long)(key & ((1u << sg.tile_shift) - 1u)))
                                          : (unsigned long long)key;
        entry_store<NR>(sg.lists, list_slot(sg, tile, pos[j]), w0, d);
        if (DET && pos[j] >= sg.cap / 2u) atomicMax(sg.fill + 2u * sg.ntiles + 2u, pos[j] + 1u);   // the host flushes before a list fills up
      } else {
        red_entry<NR>(g, key, d);       // the list is full: accumulate directly
        sg.dirty[tile] = 1u;
        if (DET) sg.fill[2u * sg.ntiles + 1u] = 1u;   // ... which is not in input order: the host reports it (GRITAS_B200_ERR_ORDER)
      }
    }
  }
}

// Sum planes of a tile per number of residual columns: nr=1 {Sx, Sxx}, nr=2 {Sx1, Sxx1, Sx2, Sxx2, Sx1x2}, nr=3 three pairs.
template <int NR>
struct TilePlanes { static constexpr int NV = (NR == 1) ? 2 : (NR == 2 ? 5 : 6); };

// ---------------------------------------------------------------------------------------------
// K3 finalize.  Output planes per box, in this order: bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs.
struct OutPlanes {
  double *bias, *rtms, *stdv, *xcov, *nobs;  // device arrays in the reference's layout (may be null)
  int xcov_planes;                           // how many ir slabs of xcov exist (1 for the raw getters)
};

// Correctly rounded x / d when several numerators share one divisor (the count m, or m-1): with r = RN(1/d),
// q = RN(x r) is a faithful quotient and one residual correction with two FMAs, q' = RN(q + RN(x - q d) r),
// is the correctly rounded quotient (Markstein's theorem; x - q d is exact in an FMA).  Non-finite or
// vanishing intermediate results fall back to the IEEE division.  Bit-identical to __ddiv_rn.
__device__ __noinline__ double div_ieee(double x, double d) { return __ddiv_rn(x, d); }   // rare: one copy, never speculated

// Square root for the statistics.  The library routine sends zero (and the other special operands) down an
// out-of-line slow path; zeros are everywhere here (empty boxes, boxes with one observation), so every warp would pay
// for it.  sqrt(+0) = +0 is answered directly, everything else is the correctly rounded library result.
__device__ __forceinline__ double sqrt_stat(double x) {
  const double s = __dsqrt_rn(x == 0.0 ? 1.0 : x);
  return x == 0.0 ? x : s;
}

__device__ __forceinline__ double div_shared(double x, double d, double r) {
  const double q = __dmul_rn(x, r);
  const double rem = __fma_rn(-q, d, x);
  const double q2 = __fma_rn(rem, r, q);
  // guard: huge / tiny magnitudes (overflow, underflow into subnormals), zero, inf and NaN are not covered by the
  // theorem (or not worth covering): one test on the exponent field, 2^-960 <= |q| < 2^960 takes the fast result
  const unsigned ex = ((unsigned)__double2hiint(q) >> 20) & 0x7ffu;
  if (ex - 63u >= 1920u) return div_ieee(x, d);
  return q2;
}

// m_gritas_grids.f:566-605 (2-D) and 612-654 (3-D) for one box.  raw: pass the sums through.
template <int NR>
__device__ __forceinline__ void finalize_box(const double *__restrict__ rec, bool is3d, int nrmlz, int ntresh,
                                             int raw, double (&bias)[NR], double (&rtms)[NR], double (&stdv)[NR],
                                             double (&xcov)[NR], double &nobs) {
  const double n = rec[0];
  nobs = n;
  if (raw == 2) {   // MinMax records: bias <- min, stdv <- max (the arrays gritas.f:770-775 passes to GrITAS_MinMax)
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) { bias[ir] = 0.0; rtms[ir] = 0.0; stdv[ir] = 0.0; xcov[ir] = 0.0; }
    bias[0] = f64_from_order_key(__double_as_longlong(rec[1]));
    stdv[0] = f64_from_order_key(__double_as_longlong(rec[2]));
    return;
  }
  if (NR == 3 && raw == 3) {   // GrITAS_3CH_Norm_ (m_gritas_grids.f:1247-1339): variances + three-cornered-hat estimates
    const int m = nrmlz ? (int)n : 1;
    const double dm = (double)m, dm1 = (double)(m - 1);
    const bool some = (n > 0.0) && (n >= (double)ntresh);
    const bool many = (n > 1.0) && (n >= (double)ntresh);
    const bool one = !many && is3d && n == 1.0;                           // :1322 (the 2-D block has no such case)
    double var[NR];
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      double tmp = 0.0;
      if (some) {
        bias[ir] = div_ieee(rec[1 + 2 * ir], dm);
        const double t = div_ieee(rec[2 + 2 * ir], dm);
        rtms[ir] = sqrt_stat(t);
        tmp = fabs(__dsub_rn(t, __dmul_rn(bias[ir], bias[ir])));
      } else {
        bias[ir] = AMISS;
        rtms[ir] = AMISS;
      }
      var[ir] = many ? div_ieee(__dmul_rn(dm, tmp), dm1) : (one ? tmp : AMISS);
      stdv[ir] = var[ir];
    }
    // the whole-array expressions of :1287-1289 / :1337-1339 run over missing values too; slots 2 and 3 are combined
    // in a different order in the 2-D and the 3-D block
    const double v1 = var[0], v2 = var[1], v3 = var[NR - 1];
    xcov[0] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v1, v2), v3));
    const double a = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v2, v3), v1)), b = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v3, v1), v2));
    xcov[1] = is3d ? b : a;
    xcov[NR - 1] = is3d ? a : b;
    return;
  }
  const double xraw = (NR == 1) ? rec[2] : (NR == 2 ? rec[5] : 0.0);  // nr=3: xcov untouched (:397-401)
  if (raw) {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) { bias[ir] = rec[1 + 2 * ir]; rtms[ir] = rec[2 + 2 * ir]; stdv[ir] = 0.0; xcov[ir] = 0.0; }
    xcov[0] = xraw;
    return;
  }
  constexpr int nx = NR;  // max(1,nr), m_gritas_grids.f:556
  const int m = nrmlz ? (int)n : 1;  // INTEGER m (m_gritas_grids.f:548, 572)
  const double dm = (double)m, dm1 = (double)(m - 1);
  const bool some = (n > 0.0) && (n >= (double)ntresh);
  const bool many = (n > 1.0) && (n >= (double)ntresh);
  // m >= 1 whenever `some` holds with nrmlz; m = 1, m-1 = 0 under -nonorm: the reciprocal path needs a finite,
  // non-zero divisor, anything else takes the plain division (inf / NaN results exactly as the reference)
  const bool fast_m = some && dm >= 1.0, fast_m1 = many && dm1 >= 1.0;
  // (the reciprocals are evaluated for every lane: keep zero divisors away from the library's slow path)
  const double rm = __drcp_rn(fast_m ? dm : 1.0), rm1 = __drcp_rn(fast_m1 ? dm1 : 1.0);
  auto div_m = [&](double x) { return fast_m ? div_shared(x, dm, rm) : div_ieee(x, dm); };
  auto div_m1 = [&](double x) { return fast_m1 ? div_shared(x, dm1, rm1) : div_ieee(x, dm1); };
  double tmp[NR], xtmp = 0.0;
  if (some) {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      bias[ir] = div_m(rec[1 + 2 * ir]);
      const double t = div_m(rec[2 + 2 * ir]);
      rtms[ir] = sqrt_stat(t);
      tmp[ir] = fabs(__dsub_rn(t, __dmul_rn(bias[ir], bias[ir])));
    }
    xtmp = div_m(xraw);
    xcov[0] = xtmp;
    xtmp = __dsub_rn(xtmp, __dmul_rn(bias[0], bias[nx - 1]));
  } else {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) { bias[ir] = AMISS; rtms[ir] = AMISS; tmp[ir] = 0.0; }
    xcov[0] = AMISS;
  }
  // many: sqrt(m*tmp/(m-1)); a single observation in a 3-D box: sqrt(tmp) (m_gritas_grids.f:642-644, the 2-D branch
  // has no such case); otherwise AMISS.  One square root per column serves both cases (a warp usually holds both).
  const bool one = !many && is3d && n == 1.0;
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) {
    const double arg = many ? div_m1(__dmul_rn(dm, tmp[ir])) : tmp[ir];
    const double sq = sqrt_stat(arg);
    stdv[ir] = (many || one) ? sq : AMISS;
  }
  xcov[nx - 1] = many ? div_m1(__dmul_rn(dm, xtmp)) : (one ? xtmp : AMISS);
}

// 3-D: tile of 32 levels x 32 longitudes for one (j,l).  Records are read level-fastest (the HBM
// layout), statistics computed in registers, all output planes transposed through shared memory in
// one go (two barriers per tile), written longitude-fastest (the reference's layout).
template <int NR>
__global__ void __launch_bounds__(1024)
k_finalize3(const GridDesc g, OutPlanes out, int nrmlz, int ntresh, int raw, int jl0, int jl1) {
  constexpr int NP = 4 * NR + 1;
  extern __shared__ double s_tile[];               // [NP][32][33]
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const size_t plane = (size_t)g.im * g.jm * g.km * g.l3d;  // one ir slab
  for (int jl = jl0 + blockIdx.z; jl < jl1; jl += gridDim.z) {
    const int j = jl % g.jm, l = jl / g.jm;
    double v[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) v[p] = 0.0;
    {
      const int k = k0 + tx, i = i0 + ty;
      if (k < g.nlevs && i < g.im) {  // levels nlevs..km-1 are never touched by the reference: zeros
        const double *rec =
            g.acc + ((size_t)k + (size_t)g.km * ((size_t)i + (size_t)g.im * ((size_t)j + (size_t)g.jm * l))) * g.rs;
        double r[8];
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(rec)), b = __ldcs(reinterpret_cast<const double2 *>(rec + 2));
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
        if (NR > 1) {
          const double2 c2 = __ldcs(reinterpret_cast<const double2 *>(rec + 4)), d2 = __ldcs(reinterpret_cast<const double2 *>(rec + 6));
          r[4] = c2.x; r[5] = c2.y; r[6] = d2.x; r[7] = d2.y;
        }
        double bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs;
        finalize_box<NR>(r, true, nrmlz, ntresh, raw, bias, rtms, stdv, xcov, nobs);
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) {
          v[ir] = bias[ir]; v[NR + ir] = rtms[ir]; v[2 * NR + ir] = stdv[ir]; v[3 * NR + ir] = xcov[ir];
        }
        v[4 * NR] = nobs;
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) s_tile[(p * 32 + ty) * 33 + tx] = v[p];
    __syncthreads();
    const int ko = k0 + ty, io = i0 + tx;
    if (ko < g.km && io < g.im) {
      const size_t o = (size_t)io + (size_t)g.im * ((size_t)j + (size_t)g.jm * ((size_t)ko + (size_t)g.km * l));
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int grp = p / NR, ir = p % NR;
        double *dst = (p == 4 * NR) ? out.nobs : (grp == 0 ? out.bias : (grp == 1 ? out.rtms : (grp == 2 ? out.stdv : out.xcov)));
        if (dst == nullptr) continue;
        if (grp == 3 && p != 4 * NR && ir >= out.xcov_planes) continue;
        __stcs(dst + o + (p == 4 * NR ? 0 : plane * ir), s_tile[(p * 32 + tx) * 33 + ty]);
      }
    }
    __syncthreads();
  }
}

// 2-D: one thread per surface box, no transpose needed.
template <int NR>
__global__ void __launch_bounds__(256)
k_finalize2(const GridDesc g, OutPlanes out, int nrmlz, int ntresh, int raw, size_t b0, size_t b1) {
  const size_t nb = g.nbox2;
  for (size_t b = b0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += (size_t)gridDim.x * blockDim.x) {
    const double *rec = g.acc + ((size_t)g.nbox3 + b) * g.rs;
    double r[8];
#pragma unroll
    for (int q = 0; q < (NR > 1 ? 8 : 4); ++q) r[q] = rec[q];
    double bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs;
    finalize_box<NR>(r, false, nrmlz, ntresh, raw, bias, rtms, stdv, xcov, nobs);
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      if (out.bias) out.bias[b + nb * ir] = bias[ir];
      if (out.rtms) out.rtms[b + nb * ir] = rtms[ir];
      if (out.stdv) out.stdv[b + nb * ir] = stdv[ir];
      if (out.xcov && ir < out.xcov_planes) out.xcov[b + nb * ir] = xcov[ir];
    }
    if (out.nobs) out.nobs[b] = nobs;
  }
}

// Adds reference-layout raw sums (bias=Sx, rtms=Sxx, xcov slot 1, nobs) into the records.
template <int NR>
__global__ void __launch_bounds__(256)
k_add_raw(const GridDesc g, OutPlanes in, int is3d) {
  const size_t nb = is3d ? g.nbox3 : g.nbox2;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    size_t recno;
    if (is3d) {
      size_t r = b;
      const size_t i = r % g.im; r /= g.im;
      const size_t j = r % g.jm; r /= g.jm;
      const size_t k = r % g.km; const size_t l = r / g.km;
      recno = k + (size_t)g.km * (i + (size_t)g.im * (j + (size_t)g.jm * l));
    } else {
      recno = (size_t)g.nbox3 + b;
    }
    double *rec = g.acc + recno * g.rs;
    if (in.nobs) rec[0] = __dadd_rn(rec[0], in.nobs[b]);
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      if (in.bias) rec[1 + 2 * ir] = __dadd_rn(rec[1 + 2 * ir], in.bias[b + nb * ir]);
      if (in.rtms) rec[2 + 2 * ir] = __dadd_rn(rec[2 + 2 * ir], in.rtms[b + nb * ir]);
    }
    if (NR == 2 && in.xcov) rec[5] = __dadd_rn(rec[5], in.xcov[b]);
  }
}

// The finalize step of a tile whose sums sit in shared memory (m_gritas_grids.f:566-654): shared by the tile kernels of the
// fast path (k_tile_final) and of the deterministic path (k_tile_det).  s_val: [NV][T] sums; the count of box b is s_nd[b]
// (IMPORTN) or s_tot[b]; addrec: the records of the box (of every rank in recmask when PEER) are added on top.
template <int NR, bool PEER, bool IMPORTN, int NTH>
__device__ __forceinline__ void tile_finalize_out(const GridDesc &g, size_t first, unsigned nb, unsigned T, const double *s_val,
                                                  const unsigned *s_tot, const double *s_nd, bool addrec, unsigned recmask,
                                                  const PeerSet &peers, const OutPlanes &o2, const OutPlanes &o3, int raw2, int raw3,
                                                  int nrmlz, int ntresh, unsigned tid) {
  constexpr int NV = TilePlanes<NR>::NV;
  const size_t plane3 = (size_t)g.nbox3, plane2 = (size_t)g.nbox2;
  auto box = [&](unsigned b, bool is3d, bool zero, int raw, const OutPlanes &out, size_t o, size_t plane) {
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = 0.0;
    if (!zero) {
      r[0] = IMPORTN ? s_nd[b] : (double)s_tot[b];
#pragma unroll
      for (int q = 0; q < NV; ++q) r[1 + q] = s_val[(size_t)q * T + b];
      if (addrec) {
        // PEER: the records of every rank that accumulated part of this tile directly, in rank order
        for (int src = 0; src < (PEER ? peers.n : 1); ++src) {
          if (PEER && !(recmask & (1u << src))) continue;
          const double *rec = (PEER ? peers.acc[src] : g.acc) + (first + b) * (size_t)g.rs;
          const double2 a0 = *reinterpret_cast<const double2 *>(rec), a1 = *reinterpret_cast<const double2 *>(rec + 2);
          r[0] = __dadd_rn(a0.x, r[0]); r[1] = __dadd_rn(a0.y, r[1]); r[2] = __dadd_rn(a1.x, r[2]);
          if (NR > 1) {
            const double2 a2 = *reinterpret_cast<const double2 *>(rec + 4), a3 = *reinterpret_cast<const double2 *>(rec + 6);
            r[3] = __dadd_rn(a1.y, r[3]); r[4] = __dadd_rn(a2.x, r[4]); r[5] = __dadd_rn(a2.y, r[5]);
            if (NR > 2) r[6] = __dadd_rn(a3.x, r[6]);
          }
        }
      }
    }
    double bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs = 0.0;
    if (zero) {
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) { bias[ir] = 0.0; rtms[ir] = 0.0; stdv[ir] = 0.0; xcov[ir] = 0.0; }
    } else {
      finalize_box<NR>(r, is3d, nrmlz, ntresh, raw, bias, rtms, stdv, xcov, nobs);
    }
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      if (out.bias) __stcs(out.bias + o + plane * ir, bias[ir]);
      if (out.rtms) __stcs(out.rtms + o + plane * ir, rtms[ir]);
      if (out.stdv) __stcs(out.stdv + o + plane * ir, stdv[ir]);
      if (out.xcov && ir < out.xcov_planes) __stcs(out.xcov + o + plane * ir, xcov[ir]);
    }
    if (out.nobs) __stcs(out.nobs + o, nobs);
  };
  const size_t end3 = min(first + nb, plane3);
  if (first < end3) {
    // the tile's span of level-fastest records covers columns c0..c1 (column = i + im*(j + jm*l)); walk it level
    // by level so that consecutive threads write consecutive longitudes
    const unsigned km = (unsigned)g.km;
    const size_t c0 = first / km, c1 = (end3 - 1) / km;
    const unsigned ncol = (unsigned)(c1 - c0 + 1);
    const unsigned head = (unsigned)(first - c0 * km), span = (unsigned)(end3 - first);
    // The span is enumerated without holes, so that a full tile is exactly T / NTH rounds: first the interior columns
    // level by level (consecutive threads -> consecutive longitudes), then the valid levels of the partial first
    // and last column.  Divisions by the column count, im and jm are multiplications: floor((x + 0.5) / d) =
    // trunc((x + 0.5) * (1/d)) as long as the rounding error stays below 0.5/d -- q < 2^16 in fp32, col < 2^32 in
    // fp64 (error < 2e-6 for any im)
    const unsigned nin = ncol > 2u ? ncol - 2u : 0u;                         // interior columns
    const unsigned nA = nin * km, nB = (ncol == 1u) ? span : km - head;     // interior cells, cells of the first column
    const float rnin = 1.0f / (float)max(nin, 1u);
    for (unsigned q = tid; q < span; q += NTH) {
      unsigned k, ci;
      if (q < nA) {
        k = (unsigned)(((float)q + 0.5f) * rnin);
        ci = 1u + (q - k * nin);
      } else if (q < nA + nB) {
        k = head + (q - nA);
        ci = 0u;
      } else {
        k = q - nA - nB;
        ci = ncol - 1u;
      }
      const unsigned rel = ci * km + k;
      if (PEER) {                                           // a tile straddling the slab boundary: the neighbour rank has the rest
        const size_t recno = first + (rel - head);
        if (recno < peers.own_lo || recno >= peers.own_hi) continue;
      }
      const unsigned col = (unsigned)c0 + ci;               // < nbox3 / km: 32 bits
      const unsigned jl = (unsigned)(((double)col + 0.5) * g.rim), i = col - jl * (unsigned)g.im;
      const unsigned l = (unsigned)(((double)jl + 0.5) * g.rjm), j = jl - l * (unsigned)g.jm;
      const size_t o = (size_t)i + (size_t)g.im * ((size_t)j + (size_t)g.jm * ((size_t)k + (size_t)km * (size_t)l));
      GB_ASSERT(rel >= head && rel - head < T && first + (rel - head) < plane3 && o < plane3 && i < (unsigned)g.im && j < (unsigned)g.jm &&
                k < km && l < (unsigned)g.l3d);
      box(rel - head, true, (int)k >= g.nlevs, raw3, o3, o, plane3);   // levels nlevs..km-1: never touched, zeros
    }
  }
  for (size_t rec = max(first, plane3) + tid; rec < first + nb; rec += NTH) {
    if (PEER && (rec < peers.own_lo || rec >= peers.own_hi)) continue;
    GB_ASSERT(rec - first < T && rec - plane3 < plane2);
    box((unsigned)(rec - first), false, false, raw2, o2, rec - plane3, plane2);
  }
}

// ---------------------------------------------------------------------------------------------
// Tiled fast path, part 2: one CTA per tile accumulates the tile's list in shared memory, then either
//   K2f (FLUSH = false): finalizes straight from shared memory -- what GrITAS_Norm runs when the month is still
//        parked in the tile lists (one process, no reduction across ranks).  Nothing is written but the output
//        planes: the records (1.7 GB at cfg2) are neither written by a flush nor read back by K3, and the lists stay
//        as they are (Norm does not change the accumulation state);
//   K2t (FLUSH = true): adds the tile to its HBM records with plain coalesced read-modify-writes (the tile is owned
//        by this CTA: no global atomics; stores when the records are known to be zero) and empties the list --
//        the flush when the lists fill up and before a reduction across ranks.
//
// No floating-point atomics: shared memory has no native fp64 add (a CAS loop per add; an earlier version of this
// kernel spent its time there).  Instead every chunk of the list is counting-sorted by box inside the CTA -- one
// native int32 shared-memory atomic per entry gives the entry its rank within the box, a scan of the per-box counts
// gives the box's start, the residuals are dropped into shared memory at start + rank -- and the thread that owns a
// box then adds the box's run with plain DADD/DFMAs into the tile's sum planes.  The planes double as the
// transposition buffer: the finalize step walks the tile level by level (the records are level-fastest, the
// reference's arrays longitude-fastest) and writes each plane in runs of ~tile/km consecutive doubles.
template <int NR>
struct TileFinal {
  static constexpr int THREADS = 512;                     // 2 or 3 CTAs per SM: their barrier-separated phases overlap
  static constexpr int EPT = (NR == 3) ? 4 : 6;           // entries per thread and chunk
  static constexpr int CH = EPT * THREADS;                // chunk of the list sorted at a time
  static constexpr int NV = TilePlanes<NR>::NV;
  // an entry's box within the tile and its rank within the box are packed into 16 + 16 bits while the chunk is sorted: tiles hold
  // at most 2^15 records (GRITAS_B200_TILE_SHIFT <= 15, checked in create) and a chunk fewer than 2^16 entries
  static_assert(CH < 65536, "rank within a chunk must fit 16 bits");
  __host__ __device__ static constexpr size_t smem_bytes(size_t T) {
    return sizeof(double) * ((size_t)NV * T + (size_t)NR * CH) + sizeof(unsigned) * (2 * T + 4 + 32);
  }
};

// MODE: what happens to the tile's sums once its list is accumulated
constexpr int TILE_FINALIZE = 0;   // K2f: statistics straight from shared memory (one rank)
constexpr int TILE_FLUSH = 1;      // K2t: add to the HBM records, empty the list
constexpr int TILE_EXPORT = 2;     // K2x: write {n, sums} (+ the tile's records, if any) as compact planes: the operand of the
                                   //      reduction across ranks -- 8*(1+NV) bytes per box instead of a 64-byte record
constexpr int TILE_IMPORT = 3;     // K3x: no list; the planes come from the reduced buffer, then the statistics as K2f
constexpr int TILE_IMPORT_PEERS = 5;   // K3p: sharded Norm, planes exchange: the tile's {n, sums} planes are summed over every rank's export buffer
                                       //      (peer memory over NVLink) and finalized -- used when the ranks park more entries than boxes
constexpr int TILE_PEERS = 4;      // K2p: K2f of a sharded multi-rank Norm -- the tile's entries are pulled from the lists of
                                   //      every rank (peer memory over NVLink), only the boxes this rank owns are finalized

#pragma nv_diag_suppress 186   // TILE_IMPORT instantiates the list loop with n = 0: 'pointless comparison' is the point
template <int NR, int MODE>
__global__ void __launch_bounds__(TileFinal<NR>::THREADS, 2)
k_tile_final(const GridDesc g, const StageDesc sg, unsigned nrec, OutPlanes o2, OutPlanes o3, int nrmlz, int ntresh,
             int raw2, int raw3, int fresh, unsigned tile0, unsigned tile1, double *__restrict__ red, const PeerSet peers) {
  constexpr bool FLUSH = MODE == TILE_FLUSH;
  constexpr bool PEER = MODE == TILE_PEERS;
  constexpr bool PIMP = MODE == TILE_IMPORT_PEERS;
  constexpr bool IMPORT = MODE == TILE_IMPORT || PIMP;
  __shared__ unsigned s_off[MAX_PEERS + 1];                // PEER: start of every rank's entries in the tile's virtual list
  __shared__ unsigned s_recmask;                           // PEER: ranks whose records hold part of this tile
  using TF = TileFinal<NR>;
  constexpr int NV = TF::NV, EPT = TF::EPT, CH = TF::CH, NTH = TF::THREADS;
  extern __shared__ __align__(16) double s_dyn[];
  const unsigned T = 1u << sg.tile_shift;
  double *s_val = s_dyn;                                   // [NV][T] sums of the tile
  double *s_stage = s_val + (size_t)NV * T;                // [NR][CH] residuals of the chunk, sorted by box
  unsigned *s_tot = reinterpret_cast<unsigned *>(s_stage + (size_t)NR * CH);   // [T] entries per box so far
  unsigned *s_cnt = s_tot + T;                             // [T+1] chunk: count per box, then exclusive start
  unsigned *s_warp = s_cnt + T + 4;                        // [32]
  const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const unsigned PT = (T + NTH - 1) / NTH;                 // counters per thread in the scan
  PeerSet pown = peers;                                    // PEER: slab boundaries balanced on the device
  if ((PEER || PIMP) && peers.bounds) {
    pown.own_lo = peers.bounds[peers.self];
    pown.own_hi = peers.bounds[peers.self + 1];
    tile0 = pown.own_lo >> sg.tile_shift;
    tile1 = (unsigned)(((size_t)pown.own_hi + T - 1u) >> sg.tile_shift);
  }
  for (unsigned tile = tile0 + blockIdx.x; tile < tile1; tile += gridDim.x) {
    if (PEER) {
      // the tile's list is the concatenation of every rank's list of this tile (a chunk of the counting sort may span
      // ranks: eight short lists cost one pass over the boxes, not eight)
      if (tid == 0u) {
        unsigned tot = 0u, mask = 0u;
        for (int s = 0; s < peers.n; ++s) {
          const unsigned *f = peers.fill[s];
          s_off[s] = tot;
          tot += min(f[tile], sg.cap);
          if (f[2u * sg.ntiles] != 0u || f[sg.ntiles + tile] != 0u) mask |= 1u << s;
        }
        s_off[peers.n] = tot;
        s_recmask = mask;
      }
      __syncthreads();
    }
    const unsigned n = IMPORT ? 0u : (PEER ? s_off[peers.n] : min(sg.fill[tile], sg.cap));
    if (FLUSH && n == 0u) continue;                       // uniform across the CTA
    // records untouched since the reset (no flush, no direct accumulate into this tile): they are zero, do not read them
    const unsigned recmask = PEER ? s_recmask : 0u;
    const bool addrec = PEER ? (recmask != 0u) : (!IMPORT && (!fresh || sg.dirty[tile] != 0u));
    // compact planes of this tile in the reduction buffer: [1 + NV][T] doubles, the counts first
    double *rplanes = (MODE == TILE_EXPORT || MODE == TILE_IMPORT) ? red + (size_t)tile * (size_t)(1 + NV) * T : nullptr;
    double *s_nd = s_stage;                                 // IMPORT: [T] counts as doubles (the staging area is idle)
    if (PIMP) {
      // every rank's planes of this tile, summed in rank order (reproducible), straight from peer memory
      const size_t toff = (size_t)tile * (size_t)(1 + NV) * T;
      for (unsigned e = tid; e < (1u + NV) * T; e += NTH) {
        double v = 0.0;
        for (int src = 0; src < peers.n; ++src) v = __dadd_rn(v, __ldcs(peers.red[src] + toff + e));
        if (e < T) s_nd[e] = v; else s_val[e - T] = v;
      }
    } else if (MODE == TILE_IMPORT) {
      for (unsigned e = tid; e < T; e += NTH) s_nd[e] = __ldcs(rplanes + e);
      for (unsigned e = tid; e < NV * T; e += NTH) s_val[e] = __ldcs(rplanes + T + e);
    } else if (n == 0u) {                                   // empty list: nothing will store the planes
      for (unsigned e = tid; e < NV * T; e += NTH) s_val[e] = 0.0;
      for (unsigned e = tid; e < T; e += NTH) s_tot[e] = 0u;
    }
    for (unsigned c0 = 0; c0 < n; c0 += CH) {
      for (unsigned e = tid; e <= T; e += NTH) s_cnt[e] = 0u;
      __syncthreads();
      // ---- rank of every entry within its box ----
      unsigned pk[EPT];
      double d[EPT][NR];
      const unsigned live_u = min((unsigned)EPT, (n - c0 + NTH - 1u) / NTH);   // uniform: rows of the chunk that hold entries
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const unsigned e = c0 + (unsigned)u * NTH + tid;
        pk[u] = KEY_NONE;
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) d[u][ir] = 0.0;
        if ((unsigned)u < live_u && e < n) {
          if (PEER) {
            int src = 0;
            while (src + 1 < peers.n && e >= s_off[src + 1]) ++src;
            pk[u] = entry_load<NR>(peers.lists[src], list_slot(sg, tile, e - s_off[src]), d[u]) & (T - 1u);
          } else {
            GB_ASSERT(tile < sg.ntiles && e < sg.cap);
            pk[u] = entry_load<NR>(sg.lists, list_slot(sg, tile, e), d[u]) & (T - 1u);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < EPT; ++u)
        if ((unsigned)u < live_u && pk[u] != KEY_NONE) pk[u] |= atomicAdd(s_cnt + pk[u], 1u) << 16;
      __syncthreads();
      // ---- exclusive scan of the counts (in place); s_cnt[T] = entries of the chunk ----
      {
        unsigned sum = 0u;
        for (unsigned q = 0; q < PT; ++q) {
          const unsigned i = tid * PT + q;
          if (i < T) sum += s_cnt[i];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31u) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0u) {
          const unsigned w = s_warp[lane];
          unsigned wi = w;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
          }
          s_warp[lane] = wi - w;
        }
        __syncthreads();
        unsigned excl = incl - sum + s_warp[wid];
        for (unsigned q = 0; q < PT; ++q) {
          const unsigned i = tid * PT + q;
          if (i < T) {
            const unsigned c = s_cnt[i];
            s_cnt[i] = excl;
            excl += c;
          }
        }
        if (tid == NTH - 1u) s_cnt[T] = excl;
      }
      __syncthreads();
      // ---- residuals to start(box) + rank ----
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        if ((unsigned)u >= live_u || pk[u] == KEY_NONE) continue;
        const unsigned pos = s_cnt[pk[u] & 0xFFFFu] + (pk[u] >> 16);
        GB_ASSERT((pk[u] & 0xFFFFu) < T && pos < (unsigned)CH && pos < s_cnt[(pk[u] & 0xFFFFu) + 1u]);
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) s_stage[(size_t)ir * CH + pos] = d[u][ir];
      }
      __syncthreads();
      // ---- the owner of a box adds its run.  The first chunk STORES the sums (zeros for empty boxes: the planes are
      // never cleared), later chunks add to them; with 512-record tiles most lists are a single chunk ----
      const bool first_chunk = c0 == 0u;
      for (unsigned b = tid; b < T; b += NTH) {
        const unsigned s0 = s_cnt[b], e0 = s_cnt[b + 1u];
        GB_ASSERT(s0 <= e0 && e0 <= (unsigned)CH);
        if (!first_chunk && e0 == s0) continue;
        double a[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) a[q] = 0.0;
#pragma unroll 1
        for (unsigned p = s0; p < e0; ++p) {                    // runs are short (~3): no unrolled prologue / main loop / tail
          double x[NR];
#pragma unroll
          for (int ir = 0; ir < NR; ++ir) x[ir] = s_stage[(size_t)ir * CH + p];
          // fused multiply-adds: the squares are not rounded separately (fast mode is order-free anyway)
          if (NR == 2) {
            a[0] = __dadd_rn(a[0], x[0]); a[1] = __fma_rn(x[0], x[0], a[1]);
            a[2] = __dadd_rn(a[2], x[NR - 1]); a[3] = __fma_rn(x[NR - 1], x[NR - 1], a[3]);
            a[4] = __fma_rn(x[0], x[NR - 1], a[4]);
          } else {
#pragma unroll
            for (int ir = 0; ir < NR; ++ir) {
              a[2 * ir] = __dadd_rn(a[2 * ir], x[ir]);
              a[2 * ir + 1] = __fma_rn(x[ir], x[ir], a[2 * ir + 1]);
            }
          }
        }
        if (first_chunk) {
#pragma unroll
          for (int q = 0; q < NV; ++q) s_val[(size_t)q * T + b] = a[q];
          s_tot[b] = e0 - s0;
        } else {
#pragma unroll
          for (int q = 0; q < NV; ++q) s_val[(size_t)q * T + b] = __dadd_rn(s_val[(size_t)q * T + b], a[q]);
          s_tot[b] += e0 - s0;
        }
      }
      // the next chunk zeroes s_cnt: every owner must be done with it
      __syncthreads();
    }
    __syncthreads();
    const size_t first = (size_t)tile << sg.tile_shift;
    const unsigned nb = (unsigned)min((size_t)T, (size_t)nrec - first);
    if (MODE == TILE_EXPORT) {
      // ---- the tile's {n, sums} as compact planes; records that already hold part of the tile are folded in ----
      for (unsigned e = tid; e < (1u + NV) * T; e += NTH) {
        const unsigned q = e >> sg.tile_shift, b = e & (T - 1u);      // plane (0: counts), box
        double v = (q == 0u) ? (double)s_tot[b] : s_val[(size_t)(q - 1u) * T + b];
        if (addrec && b < nb) v = __dadd_rn(g.acc[(first + b) * (size_t)g.rs + q], v);
        __stcs(rplanes + e, v);
      }
      __syncthreads();
      continue;
    }
    if (FLUSH) {
      // ---- add the tile to the HBM records: element e of the tile's span is (box e >> rs_shift, slot e & (rs-1));
      // each thread handles pairs of doubles (16 B), four pairs in flight.  Untouched records are zero: store, and
      // store the pad slots too, so that only whole 32-byte sectors reach L2 ----
      const int rs_shift = (g.rs == 4) ? 2 : 3;
      const bool overwrite = !addrec;
      double2 *span = reinterpret_cast<double2 *>(g.acc + first * (size_t)g.rs);
      const unsigned npair = (nb << rs_shift) >> 1;
      for (unsigned e0 = tid; e0 < npair; e0 += NTH * 4u) {
        double2 cur[4], add[4];
        bool any[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned e = e0 + u * NTH;
          any[u] = false;
          if (e < npair) {
            const unsigned el = e << 1, bxi = el >> rs_shift, slot = el & ((1u << rs_shift) - 1u);
            const unsigned cn = s_tot[bxi];
            if (cn != 0u && (overwrite || slot <= (unsigned)NV)) {
              any[u] = true;
              add[u].x = (slot == 0u) ? (double)cn : (slot <= (unsigned)NV ? s_val[(size_t)(slot - 1u) * T + bxi] : 0.0);
              add[u].y = (slot + 1u <= (unsigned)NV) ? s_val[(size_t)slot * T + bxi] : 0.0;
              cur[u] = overwrite ? make_double2(0.0, 0.0) : span[e];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (any[u]) {
            cur[u].x = __dadd_rn(cur[u].x, add[u].x);
            cur[u].y = __dadd_rn(cur[u].y, add[u].y);
            span[e0 + u * NTH] = cur[u];
          }
        }
      }
      __syncthreads();
      if (tid == 0u) sg.fill[tile] = 0u;
      continue;
    }
    // ---- finalize (m_gritas_grids.f:566-654) straight from the planes ----
    tile_finalize_out<NR, PEER || PIMP, IMPORT, TF::THREADS>(g, first, nb, T, s_val, s_tot, s_nd, addrec, recmask, pown, o2, o3, raw2, raw3,
                                                             nrmlz, ntresh, tid);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Deterministic mode on the tiled path (K2d).  Same parking of the accepted observations as the fast path, but every
// entry carries its sequence number (call order, index within the call) and the tile kernel adds the entries of a box
// IN THAT ORDER, continuing from the sums the box already has: the same sequence of fp64 additions as the reference's
// sequential loop (m_gritas_grids.f:405-474) -- bit-identical sums, whatever order the tickets were handed out in,
// however many Accum calls overlapped, and (sharded Norm, no flush in between) however many ranks parked the entries.
// Replaces the radix sort of whole files (CUB) + one-thread-per-run segment sums of round 1.
//
// Per tile: (1) count the entries per box (all ranks' lists when PEER), (2) scan, (3) cut the boxes into groups of at
// most CH entries (usually one group), (4) per group: scatter the entries to start(box) + ticket, rank every entry within
// its box by sequence number (each entry scans its box's short run), (5) the owner of a box walks its run in rank order.
// Lists longer than a chunk are re-read once per group; a single box with more than CH parked entries is summed in
// arbitrary order and reported (order_lost, GRITAS_B200_ERR_ORDER).
template <int NR>
struct TileDet {
  static constexpr int THREADS = 512;
  static constexpr int EPT = (NR == 1) ? 5 : ((NR == 3) ? 3 : 4);   // as many as leave two CTAs per SM (tiles: 2048 / 1024 / 1024 boxes)
  static constexpr int CH = EPT * THREADS;
  static constexpr int NV = TilePlanes<NR>::NV;
  __host__ __device__ static constexpr size_t smem_bytes(size_t T) {
    return sizeof(double) * ((size_t)NV * T + (size_t)NR * CH + (size_t)CH) + (size_t)CH * (2 + 2) + sizeof(unsigned) * (3 * T + 8);
  }
};
constexpr int DET_FINALIZE = 0, DET_FLUSH = 1, DET_PEERS = 2;

template <int NR, int MODE>
__global__ void __launch_bounds__(TileDet<NR>::THREADS, 2)
k_tile_det(const GridDesc g, const StageDesc sg, unsigned nrec, OutPlanes o2, OutPlanes o3, int nrmlz, int ntresh, int raw2, int raw3,
           int fresh, unsigned tile0, unsigned tile1, const PeerSet peers) {
  constexpr bool FLUSH = MODE == DET_FLUSH, PEER = MODE == DET_PEERS;
  using TD = TileDet<NR>;
  constexpr int NV = TD::NV, EPT = TD::EPT, CH = TD::CH, NTH = TD::THREADS, EB = ListEntry<NR>::BYTES;
  extern __shared__ __align__(16) double s_dyn[];
  const unsigned T = 1u << sg.tile_shift;
  double *s_val = s_dyn;                                             // [NV][T] running sums of the tile
  double *s_stage = s_val + (size_t)NV * T;                          // [NR][CH] residuals of the group at start(box) + ticket
  unsigned long long *s_seq = reinterpret_cast<unsigned long long *>(s_stage + (size_t)NR * CH);   // [CH] sequence numbers
  unsigned *s_tot = reinterpret_cast<unsigned *>(s_seq + CH);        // [T] entries per box (whole list); + record count at the end
  unsigned *s_off = s_tot + T;                                       // [T + 1] exclusive scan of s_tot
  unsigned *s_cur = s_off + T + 4;                                   // [T] tickets within the current group
  unsigned short *s_inv = reinterpret_cast<unsigned short *>(s_cur + T + 4);   // [CH] position of the entry with rank k of its box
  unsigned short *s_boxof = s_inv + CH;                              // [CH] box of the entry at a position
  __shared__ unsigned s_srcoff[MAX_PEERS + 1];
  __shared__ unsigned s_recmask, s_b1, s_warp[32];
  const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const unsigned PT = (T + NTH - 1) / NTH;
  const int nsrc = PEER ? peers.n : 1;
  PeerSet pown = peers;
  if (PEER && peers.bounds) {
    pown.own_lo = peers.bounds[peers.self];
    pown.own_hi = peers.bounds[peers.self + 1];
    tile0 = pown.own_lo >> sg.tile_shift;
    tile1 = (unsigned)(((size_t)pown.own_hi + T - 1u) >> sg.tile_shift);
  }
  for (unsigned tile = tile0 + blockIdx.x; tile < tile1; tile += gridDim.x) {
    if (tid == 0u) {
      unsigned tot = 0u, mask = 0u;
      for (int s = 0; s < nsrc; ++s) {
        const unsigned *f = PEER ? peers.fill[s] : sg.fill;
        s_srcoff[s] = tot;
        tot += min(f[tile], sg.cap);
        if (PEER && (f[2u * sg.ntiles] != 0u || f[sg.ntiles + tile] != 0u)) mask |= 1u << s;
      }
      s_srcoff[nsrc] = tot;
      s_recmask = mask;
    }
    __syncthreads();
    const unsigned n = s_srcoff[nsrc];
    if (FLUSH && (n == 0u || n < (unsigned)raw2)) { __syncthreads(); continue; }   // FLUSH: raw2 = flush only lists at least this long
    const size_t first = (size_t)tile << sg.tile_shift;
    const unsigned nb = (unsigned)min((size_t)T, (size_t)nrec - first);
    // the sums continue from this rank's records (in order: everything in them is older than anything parked); other
    // ranks' records (PEER, only after a flush there) are partial sums of their own and are added at the end
    const bool ownrec = PEER ? false : (!fresh || sg.dirty[tile] != 0u);
    const unsigned recmask = PEER ? s_recmask : 0u;
    for (unsigned e = tid; e < NV * T; e += NTH) {
      const unsigned q = e >> sg.tile_shift, b = e & (T - 1u);
      s_val[e] = (ownrec && b < nb) ? g.acc[(first + b) * (size_t)g.rs + 1u + q] : 0.0;
    }
    for (unsigned e = tid; e < T; e += NTH) s_tot[e] = 0u;
    __syncthreads();
    if (n > 0u) {
      const bool single = n <= (unsigned)CH;                          // the whole list fits one chunk: entries stay in registers
      unsigned long long w[EPT];
      unsigned tk[EPT];
      double d[EPT][NR];
      auto load = [&](unsigned e, unsigned long long &w0, double (&dd)[NR], bool full) {
        int src = 0;
        while (src + 1 < nsrc && e >= s_srcoff[src + 1]) ++src;
        const unsigned char *base = PEER ? peers.lists[src] : sg.lists;
        const size_t slot = list_slot(sg, tile, e - s_srcoff[src]);
        if (full) w0 = entry_load64<NR>(base, slot, dd);
        else w0 = entry_word0(base, slot, EB);
      };
      // ---- (1) entries per box ----
      if (single) {
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
          const unsigned e = (unsigned)u * NTH + tid;
          w[u] = ~0ull;
          if (e < n) load(e, w[u], d[u], true);
        }
#pragma unroll
        for (int u = 0; u < EPT; ++u)
          if (w[u] != ~0ull) tk[u] = atomicAdd(s_tot + ((unsigned)w[u] & (T - 1u)), 1u);   // count and ticket in one
      } else {
        for (unsigned c0 = 0; c0 < n; c0 += CH) {
#pragma unroll
          for (int u = 0; u < EPT; ++u) {
            const unsigned e = c0 + (unsigned)u * NTH + tid;
            w[u] = ~0ull;
            double dummy[NR];
            if (e < n) load(e, w[u], dummy, false);
          }
#pragma unroll
          for (int u = 0; u < EPT; ++u)
            if (w[u] != ~0ull) atomicAdd(s_tot + ((unsigned)w[u] & (T - 1u)), 1u);
        }
      }
      __syncthreads();
      // ---- (2) exclusive scan: s_off[b] = entries of the boxes before b, s_off[T] = n ----
      {
        unsigned sum = 0u;
        for (unsigned q = 0; q < PT; ++q) {
          const unsigned i = tid * PT + q;
          if (i < T) sum += s_tot[i];
        }
        unsigned incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (unsigned)o) incl += t;
        }
        if (lane == 31u) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0u) {
          const unsigned wv = s_warp[lane];
          unsigned wi = wv;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (unsigned)o) wi += t;
          }
          s_warp[lane] = wi - wv;
        }
        __syncthreads();
        unsigned excl = incl - sum + s_warp[wid];
        for (unsigned q = 0; q < PT; ++q) {
          const unsigned i = tid * PT + q;
          if (i < T) { s_off[i] = excl; excl += s_tot[i]; }
        }
        if (tid == NTH - 1u) s_off[T] = excl;
      }
      __syncthreads();
      // ---- (3) groups of boxes [b0, b1) with at most CH entries; (4), (5) per group ----
      unsigned b0 = 0u;
      while (b0 < T) {
        if (tid == 0u) {
          // largest b1 with s_off[b1] - s_off[b0] <= CH (s_off is non-decreasing)
          const unsigned lim = s_off[b0] + (unsigned)CH;
          unsigned lo = b0, hi = T;
          while (lo < hi) { const unsigned mid = (lo + hi + 1u) >> 1; if (s_off[mid] <= lim) lo = mid; else hi = mid - 1u; }
          s_b1 = (lo == b0) ? b0 + 1u : lo;                           // one box alone exceeds a chunk: taken window by window
        }
        __syncthreads();
        const unsigned b1 = s_b1;
        const unsigned gbase = s_off[b0], cnt = s_off[b1] - gbase;
        const bool giant = cnt > (unsigned)CH;
        if (giant) {
          // One box alone holds more entries than a chunk: its order cannot be kept here.  The entries are added with
          // shared-memory atomics (no order) and the host reports GRITAS_B200_ERR_ORDER; counts and sums stay complete.
          if (tid == 0u) sg.fill[2u * sg.ntiles + 1u] = 2u;
          for (unsigned c0 = 0; c0 < n; c0 += CH) {
#pragma unroll 1
            for (int u = 0; u < EPT; ++u) {
              const unsigned e = c0 + (unsigned)u * NTH + tid;
              if (e >= n) continue;
              unsigned long long wv;
              double x[NR];
              load(e, wv, x, true);
              if (((unsigned)wv & (T - 1u)) != b0) continue;
              if (NR == 2) {
                atomicAdd(s_val + 0 * T + b0, x[0]); atomicAdd(s_val + 1 * T + b0, __dmul_rn(x[0], x[0]));
                atomicAdd(s_val + 2 * T + b0, x[NR - 1]); atomicAdd(s_val + 3 * T + b0, __dmul_rn(x[NR - 1], x[NR - 1]));
                atomicAdd(s_val + 4 * T + b0, __dmul_rn(x[0], x[NR - 1]));
              } else {
#pragma unroll
                for (int ir = 0; ir < NR; ++ir) {
                  atomicAdd(s_val + (size_t)(2 * ir) * T + b0, x[ir]);
                  atomicAdd(s_val + (size_t)(2 * ir + 1) * T + b0, __dmul_rn(x[ir], x[ir]));
                }
              }
            }
          }
          __syncthreads();
          b0 = b1;
          continue;
        }
        {
          if (!single) {
            for (unsigned b = b0 + tid; b < b1; b += NTH) s_cur[b] = 0u;
            __syncthreads();
          }
          auto place = [&](unsigned long long wv, const double (&dd)[NR], unsigned ticket, bool have_ticket) {
            const unsigned box = (unsigned)wv & (T - 1u);
            if (box < b0 || box >= b1) return;
            const unsigned rel = s_off[box] - gbase + (have_ticket ? ticket : atomicAdd(s_cur + box, 1u));
            GB_ASSERT(rel < (unsigned)CH && rel < s_off[box + 1u] - gbase);
            s_seq[rel] = wv >> 16;
            s_boxof[rel] = (unsigned short)box;
#pragma unroll
            for (int ir = 0; ir < NR; ++ir) s_stage[(size_t)ir * CH + rel] = dd[ir];
          };
          if (single) {
#pragma unroll
            for (int u = 0; u < EPT; ++u)
              if (w[u] != ~0ull) place(w[u], d[u], tk[u], true);
          } else {
            for (unsigned c0 = 0; c0 < n; c0 += CH) {
              unsigned long long wv[EPT];
              double dv[EPT][NR];
#pragma unroll
              for (int u = 0; u < EPT; ++u) {
                const unsigned e = c0 + (unsigned)u * NTH + tid;
                wv[u] = ~0ull;
                if (e < n) load(e, wv[u], dv[u], true);
              }
#pragma unroll
              for (int u = 0; u < EPT; ++u)
                if (wv[u] != ~0ull) place(wv[u], dv[u], 0u, false);
            }
          }
          __syncthreads();
          const unsigned m = cnt;                                      // entries placed
          // ---- order of every box's run by sequence number: s_inv[start + k] = position of the k-th entry.  Short runs: every
          // entry counts the smaller sequence numbers of its run.  Long runs (clustered data: a radiosonde site's box holds
          // hundreds of entries) would make that quadratic: one warp sorts such a run with a bitonic network instead ----
          const unsigned R0 = sg.det_r0;
          for (unsigned pos = tid; pos < m; pos += NTH) {
            const unsigned box = s_boxof[pos];
            const unsigned st = s_off[box] - gbase, r = s_tot[box];
            if (r > R0) { s_inv[pos] = (unsigned short)pos; continue; }
            const unsigned long long mine = s_seq[pos];
            unsigned rank = 0u;
            for (unsigned q = st; q < st + r; ++q) rank += (s_seq[q] < mine) ? 1u : 0u;
            GB_ASSERT(rank < r && st + r <= (unsigned)CH);
            s_inv[st + rank] = (unsigned short)pos;
          }
          __syncthreads();
          // (the lanes of a warp look at 32 boxes at a time; the long runs among them are then sorted one after the other)
          for (unsigned bb = b0 + wid * 32u; bb < b1; bb += NTH) {
            const unsigned bl = bb + lane;
            unsigned longs = __ballot_sync(0xffffffffu, bl < b1 && s_tot[bl] > R0);
            while (longs) {
            const unsigned b = bb + (unsigned)__ffs(longs) - 1u;
            longs &= longs - 1u;
            const unsigned r = s_tot[b];
            const unsigned st = s_off[b] - gbase;
            unsigned np2 = 64u;
            while (np2 < r) np2 <<= 1;
            // bitonic sort, every compare-exchange ascending (mirror step, then half-cleaners): elements beyond r act as
            // +infinity and never move, so the run needs no padding.  Keys s_seq, payload s_inv.
            auto cmpx = [&](unsigned a, unsigned c) {
              if (c >= r) return;
              const unsigned ia = st + a, ic = st + c;
              GB_ASSERT(a < c && ic < (unsigned)CH);
              const unsigned long long ka = s_seq[ia], kc = s_seq[ic];
              if (ka > kc) {
                s_seq[ia] = kc;
                s_seq[ic] = ka;
                const unsigned short t = s_inv[ia];
                s_inv[ia] = s_inv[ic];
                s_inv[ic] = t;
              }
            };
            for (unsigned kk = 2u; kk <= np2; kk <<= 1) {
              const unsigned hk = kk >> 1;
              for (unsigned i = lane; i < (np2 >> 1); i += 32u) {
                const unsigned blk = i / hk, off = i - blk * hk;
                cmpx(blk * kk + off, blk * kk + kk - 1u - off);
              }
              __syncwarp();
              for (unsigned j = kk >> 2; j >= 1u; j >>= 1) {
                for (unsigned i = lane; i < (np2 >> 1); i += 32u) {
                  const unsigned a = (i / j) * 2u * j + (i % j);
                  cmpx(a, a + j);
                }
                __syncwarp();
              }
            }
            }
          }
          __syncthreads();
          // ---- the owner of a box adds its run in sequence order, continuing the box's sums (one multiply and one add
          // per term, as the reference: no fused multiply-add) ----
          for (unsigned b = b0 + tid; b < b1; b += NTH) {
            const unsigned st = s_off[b] - gbase, r = s_tot[b];
            if (r == 0u) continue;
            double a[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) a[q] = s_val[(size_t)q * T + b];
#pragma unroll 1
            for (unsigned k = 0; k < r; ++k) {
              const unsigned pos = s_inv[st + k];
              GB_ASSERT(st + r <= (unsigned)CH && pos < (unsigned)CH && pos >= st && pos < st + r);
              double x[NR];
#pragma unroll
              for (int ir = 0; ir < NR; ++ir) x[ir] = s_stage[(size_t)ir * CH + pos];
              if (NR == 2) {
                a[0] = __dadd_rn(a[0], x[0]); a[1] = __dadd_rn(a[1], __dmul_rn(x[0], x[0]));
                a[2] = __dadd_rn(a[2], x[NR - 1]); a[3] = __dadd_rn(a[3], __dmul_rn(x[NR - 1], x[NR - 1]));
                a[4] = __dadd_rn(a[4], __dmul_rn(x[0], x[NR - 1]));
              } else {
#pragma unroll
                for (int ir = 0; ir < NR; ++ir) {
                  a[2 * ir] = __dadd_rn(a[2 * ir], x[ir]);
                  a[2 * ir + 1] = __dadd_rn(a[2 * ir + 1], __dmul_rn(x[ir], x[ir]));
                }
              }
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) s_val[(size_t)q * T + b] = a[q];
          }
          __syncthreads();
        }
        __syncthreads();
        b0 = b1;
      }
    }
    // counts: parked entries + what this rank's records already held (an integer below 2^32)
    if (ownrec) {
      for (unsigned b = tid; b < nb; b += NTH) s_tot[b] += (unsigned)g.acc[(first + b) * (size_t)g.rs];
      __syncthreads();
    }
    if (FLUSH) {
      for (unsigned b = tid; b < nb; b += NTH) {
        double *rec = g.acc + (first + b) * (size_t)g.rs;
        if (s_tot[b] == 0u && !ownrec) continue;                       // untouched records are zero already
        rec[0] = (double)s_tot[b];
#pragma unroll
        for (int q = 0; q < NV; ++q) rec[1 + q] = s_val[(size_t)q * T + b];
      }
      __syncthreads();
      if (tid == 0u) { sg.fill[tile] = 0u; sg.dirty[tile] = 1u; }      // the tile's records now hold sums
      __syncthreads();
      continue;
    }
    tile_finalize_out<NR, PEER, false, NTH>(g, first, nb, T, s_val, s_tot, nullptr, PEER && recmask != 0u, recmask, pown, o2, o3, raw2, raw3,
                                            nrmlz, ntresh, tid);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Pre-sort (gritas.f:499-508): IndexSet + eight successive stable ascending index sorts.  One LSD pass = gather the
// pass's column through the current index into an order-preserving unsigned key, then a stable radix sort of
// (key, index) pairs.
__global__ void __launch_bounds__(256)
k_iota(int *__restrict__ idx, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) idx[i] = i;
}
__global__ void __launch_bounds__(256)
k_sortkey_i32(const int *__restrict__ col, const int *__restrict__ idx, unsigned *__restrict__ key, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    key[i] = (unsigned)col[idx[i]] ^ 0x80000000u;                       // signed order -> unsigned order
}
__global__ void __launch_bounds__(256)
k_sortkey_f64(const double *__restrict__ col, const int *__restrict__ idx, unsigned long long *__restrict__ key, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    // -0.0 and +0.0 compare equal in the reference's merge (strict <): canonicalise before taking the bit pattern
    const double x = __dadd_rn(col[idx[i]], 0.0);
    key[i] = (unsigned long long)f64_order_key(x) ^ 0x8000000000000000ull;
  }
}
// -meanres (gritas.f:653-664, 665-677, 680-691): an observation whose value differs between the member and the ensemble-mean
// ODS is made passive before the filter runs (qcexcl = X_PASSIVE).
__global__ void __launch_bounds__(256)
k_meanres_qc(const double *__restrict__ obs, const double *__restrict__ mean_obs, const int *__restrict__ qcexcl, int x_passive,
             int *__restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = (obs[i] == mean_obs[i]) ? qcexcl[i] : x_passive;
}

// ob_flag as GrITAS_Accum_ leaves it (m_gritas_grids.f:406-416), from the three integer columns alone: an observation
// that came in flagged keeps its flag, one whose (kx, kt) is not in the table gets 1 -- the same value every accumulate
// kernel writes (classify4: flag_after).  Run as soon as kx / kt / ob_flag have reached the device, so that the
// write-back of ob_flag to the caller overlaps the upload of the fp64 columns (host callers without a land mask).
__global__ void __launch_bounds__(256)
k_early_flags(const GridDesc g, const int *__restrict__ kx, const int *__restrict__ kt, const int *__restrict__ flag_in,
              int *__restrict__ flag_out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int fl = flag_in[i];
    if (fl == 0) {
      const unsigned kx1 = (unsigned)(kx[i] - 1), kt1 = (unsigned)(kt[i] - 1);
      const bool inr = kx1 < (unsigned)g.kxmax && kt1 < (unsigned)g.ktmax;
      const int l = inr ? __ldg(g.table + (kx1 + (unsigned)g.kxmax * kt1)) : 0;
      if (l == 0) fl = 1;
    }
    flag_out[i] = fl;
  }
}

// The column gathers of gritas.f:510-523: dst(i) = src(is(i)) for up to 8 fp64 and 5 int32 columns in one pass.
struct GatherCols {
  const double *f_src[8]; double *f_dst[8];
  const int *i_src[5]; int *i_dst[5];
  int nf, ni;
};
__global__ void __launch_bounds__(256)
k_gather_cols(const GatherCols c, const int *__restrict__ is, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int j = is[i];
    for (int q = 0; q < c.nf; ++q) c.f_dst[q][i] = c.f_src[q][j];
    for (int q = 0; q < c.ni; ++q) c.i_dst[q][i] = c.i_src[q][j];
  }
}
__global__ void __launch_bounds__(256)
k_add_base(const int *__restrict__ idx, int *__restrict__ out, int base, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = idx[i] + base;
}

#pragma nv_diag_default 186

// Reset when nothing was flushed into the records since the last reset (everything parked in the tile lists or
// consumed by K2f): only tiles that took a direct accumulate (bin / list overflow, sg.dirty) hold anything.
__global__ void __launch_bounds__(256)
k_zero_dirty_tiles(double *__restrict__ acc, size_t ndouble, int rs, const StageDesc sg) {
  for (unsigned tile = blockIdx.x; tile < sg.ntiles; tile += gridDim.x) {
    if (sg.dirty[tile] == 0u) continue;
    const size_t first = ((size_t)tile << sg.tile_shift) * (size_t)rs;
    const size_t n = min(((size_t)1 << sg.tile_shift) * (size_t)rs, ndouble - first);
    double2 *a2 = reinterpret_cast<double2 *>(acc + first);
    for (size_t e = threadIdx.x; e < (n >> 1); e += blockDim.x) a2[e] = make_double2(0.0, 0.0);
  }
}

// Scorecard (m_gritas_grids.f:1671-1826, scores2_ + ave_nomiss): regional averages of the finalized mean / rms / stdv
// grids over the boxes that are not AMISS.  One thread per (region, level, variable, statistic) walks its region in the
// reference's order (latitude outer, longitude inner) and sums in fp64 exactly as ave_nomiss does; the result is the
// reference's real(4) collect_stats(mregs, km, l2d + l3d, nevals) entry.  Variables: the l2d surface grids first (level
// slot 1 only), then the l3d upper-air grids (:1718-1731).
struct Planes3 { const double *p[3]; };    // mean, rms, stdv
__global__ void __launch_bounds__(128)
k_scores(const GridDesc g, const Planes3 s2, const Planes3 s3, const double *__restrict__ nobs_2d,
         const double *__restrict__ nobs_3d, const int *__restrict__ regindx, int nregs, int lon0, int lon1,
         float *__restrict__ collect) {
  const int nvar = g.l2d + g.l3d, total = nregs * g.km * nvar * 3;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int mr = e % nregs, kk = (e / nregs) % g.km, nv = (e / (nregs * g.km)) % nvar, ne = e / (nregs * g.km * nvar);
    const bool surf = nv < g.l2d;
    if (surf && kk != 0) continue;                          // collect_stats(mr,1,nv,ne) only; the other levels stay as they were
    const size_t plane = (size_t)g.im * g.jm;
    const double *x = surf ? s2.p[ne] + plane * nv : s3.p[ne] + plane * ((size_t)kk + (size_t)g.km * (nv - g.l2d));
    const double *nb = surf ? nobs_2d + plane * nv : nobs_3d + plane * ((size_t)kk + (size_t)g.km * (nv - g.l2d));
    double add = -AMISS, counter = 0.0;                     // :1806-1807
    const int j0 = regindx[2 * mr], j1 = regindx[2 * mr + 1];
    for (int jj = j0; jj <= j1; ++jj)
      for (int ii = lon0; ii <= lon1; ++ii) {
        const size_t o = (size_t)(ii - 1) + (size_t)g.im * (size_t)(jj - 1);
        const double v = x[o];
        if (nb[o] >= 0.0 && fabs(__dsub_rn(v, AMISS)) > 0.01) {   // threshold = 0 (:1803), :1810
          add = (add == -AMISS) ? v : __dadd_rn(add, v);
          counter = __dadd_rn(counter, 1.0);
        }
      }
    collect[e] = (float)((add == -AMISS) ? AMISS : __ddiv_rn(add, counter));   // :1820-1824, stored as real(4)
  }
}

// Sharded Norm: local snapshot of every rank's [tickets | dirty | records-touched] words (one launch instead of a
// peer copy per rank; the tile kernel then reads them from local memory).
__global__ void __launch_bounds__(256)
k_gather_words(unsigned *__restrict__ dst, const PeerSet src, unsigned nwords) {
  for (int s = 0; s < src.n; ++s)
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < nwords; e += gridDim.x * blockDim.x)
      dst[(size_t)s * nwords + e] = src.fill[s][e];
}

// Sharded Norm: slab boundaries that give every rank about the same number of parked entries (the tile kernel's time
// follows the entries, and rows near the equator / frequent variables hold many more than polar rows).  Every rank runs this
// on the same snapshot of all ranks' fill counts, so every rank gets the same boundaries.  Boundaries are whole rows
// ((j,l) rows of im*km records in the 3-D part, of im records in the 2-D part).  One block.
__global__ void __launch_bounds__(1024)
k_slab_bounds(const unsigned *__restrict__ fills, unsigned nwords, int nranks, StageDesc sg, unsigned nrec, unsigned nbox3, unsigned row3,
              unsigned row2, unsigned long long *__restrict__ prefix, unsigned *__restrict__ bounds) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_base;
  const unsigned tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  if (tid == 0u) s_base = 0ull;
  __syncthreads();
  // weight of a tile: its parked entries on all ranks, plus one per box for the finalize work it carries anyway
  for (unsigned o = 0; o < sg.ntiles; o += 1024u) {
    const unsigned t = o + tid;
    unsigned long long v = 0ull;
    if (t < sg.ntiles) {
      for (int r = 0; r < nranks; ++r) v += min(fills[(size_t)r * nwords + t], sg.cap);
      v += 1ull << sg.tile_shift;
    }
    unsigned long long inc = v;
#pragma unroll
    for (int sft = 1; sft < 32; sft <<= 1) {
      const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, sft);
      if (lane >= (unsigned)sft) inc += x;
    }
    if (lane == 31u) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0u) {
      const unsigned long long w = s_w[lane];
      unsigned long long wi = w;
#pragma unroll
      for (int sft = 1; sft < 32; sft <<= 1) {
        const unsigned long long x = __shfl_up_sync(0xffffffffu, wi, sft);
        if (lane >= (unsigned)sft) wi += x;
      }
      s_w[lane] = wi - w;
    }
    __syncthreads();
    const unsigned long long excl = s_base + s_w[wid] + inc - v;
    if (t < sg.ntiles) prefix[t] = excl;
    __syncthreads();
    if (tid == 1023u) s_base = excl + v;
    __syncthreads();
  }
  if (tid == 0u) prefix[sg.ntiles] = s_base;
  __syncthreads();
  if (tid <= (unsigned)nranks) {
    unsigned b;
    if (tid == 0u) b = 0u;
    else if (tid == (unsigned)nranks) b = nrec;
    else {
      const unsigned long long total = prefix[sg.ntiles], target = total / (unsigned)nranks * tid + total % (unsigned)nranks * tid / (unsigned)nranks;
      unsigned lo = 0u, hi = sg.ntiles;                       // last tile whose prefix <= target
      while (lo < hi) { const unsigned mid = (lo + hi + 1u) >> 1; if (prefix[mid] <= target) lo = mid; else hi = mid - 1u; }
      const unsigned T = 1u << sg.tile_shift;
      const unsigned long long wt = prefix[min(lo + 1u, sg.ntiles)] - prefix[lo];
      unsigned long long rec = (unsigned long long)lo * T + (wt ? (target - prefix[lo]) * T / wt : 0ull);
      if (rec > nrec) rec = nrec;
      if (rec <= nbox3) rec = (rec + row3 / 2u) / row3 * row3;                      // nearest row of the 3-D part
      else rec = nbox3 + (rec - nbox3 + row2 / 2u) / row2 * row2;
      b = (unsigned)min(rec, (unsigned long long)nrec);
    }
    bounds[tid] = b;
  }
  __syncthreads();
  if (tid == 0u) {                                            // monotone whatever the rounding did
    for (int r = 1; r <= nranks; ++r) bounds[r] = max(bounds[r], bounds[r - 1]);
    const unsigned long long entries = prefix[sg.ntiles] - ((unsigned long long)sg.ntiles << sg.tile_shift);
    bounds[nranks + 1] = (unsigned)min(entries, 0xFFFFFFFFull);   // parked entries of all ranks: entry pull or planes exchange?
  }
}

// Sharded Norm on the direct / deterministic path (no lists): the records of rows this rank owns are summed over the
// ranks in rank order (reproducible) into this rank's records; K3 then finalizes them.  One-shot, like the reduction it replaces.
__global__ void __launch_bounds__(256)
k_sum_peer_records(double *__restrict__ acc, const PeerSet peers, int self, size_t d0, size_t d1) {
  for (size_t e = d0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < d1; e += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < peers.n; ++s) v = __dadd_rn(v, (s == self ? acc : peers.acc[s])[e]);
    acc[e] = v;
  }
}

}  // namespace gb200
