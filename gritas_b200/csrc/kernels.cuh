// kernels.cuh -- sm_100a device code of the GrITAS gridding hot path.
//
// K1  box index + QC filter      reference m_gritas_grids.f:405-461, gritas.f:562-587, 710-739
// K2a fast accumulate            reference m_gritas_grids.f:432-441, 465-470 (warp-aggregated fp64 RED)
// K2b deterministic accumulate   same lines; stable sort by box key, then in-order segment sums
// K3  finalize                   reference m_gritas_grids.f:566-654
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

namespace gb200 {

constexpr unsigned KEY_NONE = 0xFFFFFFFFu;
constexpr unsigned long long BAD_NONE = ~0ull;
constexpr double AMISS = 1.0e15;  // m_gritas_grids.f:72
constexpr int ACC_THREADS = 256;
constexpr int OBS_PER_THREAD = 4;

struct Status {
  unsigned long long first_bad;  // (call_seq << 32) | obs index of the first level miss, or BAD_NONE
  double bad_lev;
  unsigned ticket;
  unsigned pad;
};

struct GridDesc {
  int im, jm, km, nlevs, l2d, l3d, nr, rs;
  int kxmax, ktmax;
  int contiguous;  // p1(k+1)==p2(k), strictly descending: branch-free lower bound is exact
  int p2_in_smem;
  double dlon, dlat, p1_top;
  unsigned nbox3, nbox2;
  const int *table;
  const double *p1, *p2;
  double *acc;
};

struct ObsCols {
  const double *lat, *lon, *lev;
  const double *d[3];        // residual source columns (del(:,ir), or omf / oma)
  const double *obs, *xvec, *omf;  // only for scale_res
  const int *kx, *kt;
  const int *flag;           // ob_flag (plain entry) or qcexcl (ODS entry); may be null (all 0)
  const int *time;           // may be null
  int *flag_out;             // may be null
  int n;
  int aligned;               // every non-null column pointer is 16-byte aligned
};

struct FilterDesc {
  int ods;        // 1: `flag` is qcexcl and the gritas.f:710-739 filter is applied
  int ioptn, t1, t2, use_time, use_lon;
  double lona, lonb;
  int x_passive, x_pre_bad;
  int scale_res;  // gritas.f:564-572
};

// Staging area of the tiled fast path.  The record grid is cut into tiles of
// 2^tile_shift consecutive records; K1 appends every accepted observation as a compact
// (record key, residuals) entry to the list of its tile; K2t later accumulates one tile at a time in
// shared memory.  Lists have a fixed capacity; entries that do not fit are accumulated directly.
struct StageDesc {
  // level 2: one list of compact entries per tile of 2^tile_shift records
  unsigned *fill;      // [ntiles (padded)] entries appended since the last flush (may exceed cap: clamp)
  unsigned *skey;      // [ntiles * cap]
  double *sd[3];       // [ntiles * cap] per residual column
  unsigned cap, ntiles;
  int tile_shift;
  // level 1: per-bucket scratch of the current Accum call (a bucket = 2^(bucket_shift-tile_shift) tiles);
  // small enough to stay in L2 between the two partition kernels
  unsigned *fill1, *fill1_next;  // [PB_MAXB] ping-pong: the kernel of call n zeroes the counters of call n+1
  unsigned *key1;
  double *d1[3];
  unsigned cap1, nbuckets;
  int bucket_shift;
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

// gritas.f:710-739.  Returns the ob_flag the reference would hand to GrITAS_Accum_.
__device__ __forceinline__ int qc_filter(const FilterDesc &f, int qcexcl, int time, double lon) {
  const bool passive = (qcexcl == f.x_passive) | (qcexcl == f.x_pre_bad);
  const bool want = (f.ioptn > 0) & passive;
  const bool tsel = !f.use_time | ((f.t1 <= time) & (time <= f.t2));
  const bool lsel = !f.use_lon | ((f.lona <= lon) & (lon <= f.lonb));
  return (tsel & lsel & ((qcexcl == 0) | want)) ? 0 : 1;
}

// m_gritas_grids.f:450-456.  0-based level, or -1 ("could not find level").
__device__ __forceinline__ int find_level(const GridDesc &g, const double *__restrict__ p2s, double lev) {
  const int nl = g.nlevs;
  if (g.contiguous) {
    // level = #{kk < nl-1 : lev <= p2(kk)}; p2 strictly descending so the predicate holds on a prefix.
    int pos = 0;
    const int len = nl - 1;
    int step = 1;
    while ((step << 1) <= len) step <<= 1;
    for (; step > 0; step >>= 1) {
      const int np = pos + step;
      if (np <= len && lev <= p2s[np - 1]) pos = np;
    }
    const double top = (pos == 0) ? g.p1_top : p2s[pos - 1];
    return (lev <= top && lev > p2s[pos]) ? pos : -1;
  }
  for (int kk = 0; kk < nl; ++kk)
    if (lev <= __ldg(g.p1 + kk) && lev > __ldg(g.p2 + kk)) return kk;
  return -1;
}

// m_gritas_grids.f:409-426 + level search.  outcome: 0 accumulate, 1 not in table, 2 level miss.
__device__ __forceinline__ unsigned box_key(const GridDesc &g, const double *__restrict__ p2s, double lat,
                                            double lon, double lev, int kx, int kt, int &outcome) {
  int l = 0;
  if (kx >= 1 && kx <= g.kxmax && kt >= 1 && kt <= g.ktmax)
    l = __ldg(g.table + (size_t)(kx - 1) + (size_t)g.kxmax * (size_t)(kt - 1));
  const bool surface = l < 0;
  l = abs(l);
  if (l <= 0) { outcome = 1; return KEY_NONE; }
  long long ii = 1 + nint_sat(__ddiv_rn(__dsub_rn(lon, -180.0), g.dlon));
  if (ii == (long long)g.im + 1) ii = 1;
  if (ii == 0) ii = g.im;
  long long jj = 1 + nint_sat(__ddiv_rn(__dsub_rn(lat, -90.0), g.dlat));
  ii = min((long long)g.im, max(1ll, ii));
  jj = min((long long)g.jm, max(1ll, jj));
  const unsigned i = (unsigned)(ii - 1), j = (unsigned)(jj - 1);
  outcome = 0;
  if (surface) return g.nbox3 + i + (unsigned)g.im * (j + (unsigned)g.jm * (unsigned)(l - 1));
  const int k = find_level(g, p2s, lev);
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

__device__ __forceinline__ void report_miss(Status *st, unsigned call_seq, int n) {
  atomicMin(&st->first_bad, ((unsigned long long)call_seq << 32) | (unsigned)n);
}

// Last block to finish captures the offending level (what the reference prints before exit(7)).
__device__ __forceinline__ void finish_status(Status *st, unsigned call_seq, const double *lev) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&st->ticket, 1u);
    if (t == gridDim.x - 1) {
      st->ticket = 0;
      const unsigned long long fb = atomicAdd(&st->first_bad, 0ull);
      if (fb != BAD_NONE && (unsigned)(fb >> 32) == call_seq) st->bad_lev = lev[(unsigned)fb];
      __threadfence();
    }
  }
}

// streaming 128-bit loads that do not pollute L1
__device__ __forceinline__ double2 ldg_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream4(const int *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <int NR>
struct Obs4 {
  double lat[4], lon[4], lev[4], d[NR][4];
  int kx[4], kt[4], fl[4], tm[4];
};

template <int NR>
__device__ __forceinline__ void load_obs4(const ObsCols &c, const FilterDesc &f, int base, Obs4<NR> &o) {
  const int n = c.n;
  if (c.aligned && base + 3 < n) {
    double2 a, b;
    a = ldg_stream2(c.lat + base); b = ldg_stream2(c.lat + base + 2);
    o.lat[0] = a.x; o.lat[1] = a.y; o.lat[2] = b.x; o.lat[3] = b.y;
    a = ldg_stream2(c.lon + base); b = ldg_stream2(c.lon + base + 2);
    o.lon[0] = a.x; o.lon[1] = a.y; o.lon[2] = b.x; o.lon[3] = b.y;
    a = ldg_stream2(c.lev + base); b = ldg_stream2(c.lev + base + 2);
    o.lev[0] = a.x; o.lev[1] = a.y; o.lev[2] = b.x; o.lev[3] = b.y;
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      a = ldg_stream2(c.d[ir] + base); b = ldg_stream2(c.d[ir] + base + 2);
      o.d[ir][0] = a.x; o.d[ir][1] = a.y; o.d[ir][2] = b.x; o.d[ir][3] = b.y;
    }
    int4 v = ldg_stream4(c.kx + base);
    o.kx[0] = v.x; o.kx[1] = v.y; o.kx[2] = v.z; o.kx[3] = v.w;
    v = ldg_stream4(c.kt + base);
    o.kt[0] = v.x; o.kt[1] = v.y; o.kt[2] = v.z; o.kt[3] = v.w;
    if (c.flag) {
      v = ldg_stream4(c.flag + base);
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
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = base + j;
      const bool in = i < n;
      const int ii = in ? i : 0;
      o.lat[j] = in ? __ldg(c.lat + ii) : 0.0;
      o.lon[j] = in ? __ldg(c.lon + ii) : 0.0;
      o.lev[j] = in ? __ldg(c.lev + ii) : 0.0;
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) o.d[ir][j] = in ? __ldg(c.d[ir] + ii) : 0.0;
      o.kx[j] = in ? __ldg(c.kx + ii) : 0;
      o.kt[j] = in ? __ldg(c.kt + ii) : 0;
      o.fl[j] = (in && c.flag) ? __ldg(c.flag + ii) : 0;
      o.tm[j] = (in && f.use_time) ? __ldg(c.time + ii) : 0;
    }
  }
}

// Per-observation front half shared by the fast kernel and the key kernel: filter, table lookup,
// box index.  Returns the record key (KEY_NONE if nothing is to be accumulated) and the ob_flag
// the reference would leave behind.
__device__ __forceinline__ unsigned classify(const GridDesc &g, const FilterDesc &f, const double *p2s,
                                             double lat, double lon, double lev, int kx, int kt, int flag_in,
                                             int tm, int n, Status *st, unsigned call_seq, int &flag_after) {
  int fl = flag_in;
  if (f.ods) fl = qc_filter(f, flag_in, tm, lon);
  flag_after = fl;
  if (fl != 0) return KEY_NONE;  // m_gritas_grids.f:406-407
  int outcome;
  const unsigned key = box_key(g, p2s, lat, lon, lev, kx, kt, outcome);
  if (outcome == 1) flag_after = 1;  // m_gritas_grids.f:413-416
  if (outcome == 2) report_miss(st, call_seq, n);
  return key;
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

__device__ __forceinline__ void load_p2(const GridDesc &g, double *s_p2) {
  if (g.p2_in_smem)
    for (int k = threadIdx.x; k < g.nlevs; k += blockDim.x) s_p2[k] = g.p2[k];
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// K1+K2a fused: stream the columns, classify, accumulate with fp64 RED into the record grid.
// AGG: lanes of a warp that hit the same box (__match_any_sync) are summed in registers first and
// only the lowest lane issues the atomics.
template <int NR, bool AGG>
__global__ void __launch_bounds__(ACC_THREADS)
k_accum_fast(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq) {
  extern __shared__ double s_p2[];
  load_p2(g, s_p2);
  const double *p2s = g.p2_in_smem ? s_p2 : g.p2;
  const int nvec = (c.n + OBS_PER_THREAD - 1) / OBS_PER_THREAD;
  const int nvec_w = (nvec + 31) & ~31;  // whole warps iterate together (match/shfl need all lanes)
  const int lane = threadIdx.x & 31;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec_w; v += gridDim.x * blockDim.x) {
    const int base = v * OBS_PER_THREAD;
    Obs4<NR> o;
    if (v < nvec) load_obs4<NR>(c, f, base, o);
    int fa[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned key = KEY_NONE;
      fa[j] = 0;
      if (v < nvec && base + j < c.n)
        key = classify(g, f, p2s, o.lat[j], o.lon[j], o.lev[j], o.kx[j], o.kt[j], o.fl[j], o.tm[j], base + j, st,
                       call_seq, fa[j]);
      double sx[NR], sq[NR], sxy = 0.0;
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) {
        double d = o.d[ir][j];
        if (f.scale_res && key != KEY_NONE) d = scale_residual(f, c, d, base + j);
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
  }
  finish_status(st, call_seq, c.lev);
}

// ---------------------------------------------------------------------------------------------
// K1 alone (deterministic path): keys + identity permutation for the stable sort.
__global__ void __launch_bounds__(ACC_THREADS)
k_box_keys(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq,
           unsigned *__restrict__ keys, unsigned *__restrict__ idx, unsigned nrec) {
  extern __shared__ double s_p2[];
  load_p2(g, s_p2);
  const double *p2s = g.p2_in_smem ? s_p2 : g.p2;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < c.n; n += gridDim.x * blockDim.x) {
    const int fl = c.flag ? __ldg(c.flag + n) : 0;
    const int tm = f.use_time ? __ldg(c.time + n) : 0;
    int fa;
    const unsigned key = classify(g, f, p2s, __ldg(c.lat + n), __ldg(c.lon + n), __ldg(c.lev + n), __ldg(c.kx + n),
                                  __ldg(c.kt + n), fl, tm, n, st, call_seq, fa);
    keys[n] = key < nrec ? key : nrec;  // rejected observations sort last within key_bits
    idx[n] = (unsigned)n;
    if (c.flag_out) c.flag_out[n] = fa;
  }
  finish_status(st, call_seq, c.lev);
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
      d[ir] = scale_residual(f, c, __ldg(c.d[ir] + i), i);
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
// Tiled fast path, part 1: two-level partition of the accepted observations into per-tile lists.
//
// Scattered 4/8-byte stores run at ~40 G/s on B200 (profiles/r1_microbench_tickets_scatter.txt): far
// below what this path needs.  Both partition kernels therefore group a batch of entries by
// destination in shared memory and write every destination's entries as one run whose start and
// length are multiples of 8 entries (32 B of keys, 64 B of residuals): only full, aligned sectors
// reach L2.  Runs are padded with KEY_NONE sentinels; readers skip them.
//   K1p  k_part_obs    columns -> classify (filter, table, box index) -> per-BUCKET scratch lists
//   K1q  k_part_tiles  bucket scratch (still in L2) -> per-TILE lists in HBM
constexpr int PB_MAXB = 128;                      // destinations per grouping step
constexpr unsigned PB_OVERFLOW = 0xFFFFFFFFu;
// K1p: 256 threads x 8 observations, 3 CTAs per SM so that load / group / write-out phases of
// different CTAs overlap.  K1q: 1024 threads x 4 entries (entries only, few registers), 2 CTAs per SM.
constexpr int P1_THREADS = 256, P1_ITEMS = 8, P1_BATCH = P1_THREADS * P1_ITEMS, P1_SCAP = P1_BATCH + 8 * PB_MAXB;
constexpr int P2_THREADS = 1024, P2_ITEMS = 4, P2_BATCH = P2_THREADS * P2_ITEMS, P2_SCAP = P2_BATCH + 8 * PB_MAXB;

template <int NR>
struct PartView {
  double *d[NR];        // [SCAP] grouped residuals
  unsigned *key;        // [SCAP] grouped keys
  unsigned *cnt;        // [PB_MAXB] entries per destination in this batch
  unsigned *off;        // [PB_MAXB] start of the destination's (padded) group in the staging arrays
  unsigned *run;        // [PB_MAXB] start of the run inside the destination's global list, or PB_OVERFLOW
  unsigned *total;      // padded entries in this batch
  unsigned char *grp;   // [SCAP/8] destination of each group of 8 staged entries
};

template <int NR, int SCAP>
__host__ __device__ constexpr size_t part_smem_bytes() {
  return sizeof(double) * NR * SCAP + sizeof(unsigned) * (SCAP + 3 * PB_MAXB + 4) + SCAP / 8;
}

template <int NR, int SCAP>
__device__ __forceinline__ PartView<NR> part_view(unsigned char *smem) {
  PartView<NR> v;
  double *dp = reinterpret_cast<double *>(smem);
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) v.d[ir] = dp + (size_t)ir * SCAP;
  unsigned *up = reinterpret_cast<unsigned *>(dp + (size_t)NR * SCAP);
  v.key = up; up += SCAP;
  v.cnt = up; up += PB_MAXB;
  v.off = up; up += PB_MAXB;
  v.run = up; up += PB_MAXB;
  v.total = up; up += 4;
  v.grp = reinterpret_cast<unsigned char *>(up);
  return v;
}

// After the counting pass: lay the destinations out back to back (each padded to a multiple of 8),
// reserve the same number of entries at the tail of every destination's global list, write the
// pad sentinels and the group->destination map.  Called by all threads of the CTA.
template <int NR>
__device__ __forceinline__ void part_layout(const PartView<NR> &v, int nb, unsigned *__restrict__ g_fill,
                                            unsigned *__restrict__ g_key, unsigned cap, unsigned list_first) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    unsigned p[4], local = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int b = lane * 4 + q;
      const unsigned c = b < nb ? v.cnt[b] : 0u;
      p[q] = (c + 7u) & ~7u;
      local += p[q];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    unsigned at = incl - local;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      v.off[lane * 4 + q] = at;
      at += p[q];
    }
    if (lane == 31) *v.total = incl;
  }
  __syncthreads();
  if (tid < nb) {
    const unsigned c = v.cnt[tid], padded = (c + 7u) & ~7u, off = v.off[tid];
    unsigned run = 0;
    if (c) {
      const unsigned gpos = atomicAdd(g_fill + list_first + tid, padded);
      if (gpos + padded <= cap) {
        run = gpos;
      } else {
        // does not fit: these entries are accumulated directly by the caller.  Exactly one reservation
        // straddles the capacity; it closes the unwritten gap with sentinels so readers never see garbage.
        run = PB_OVERFLOW;
        unsigned *lk = g_key + (size_t)(list_first + tid) * cap;
        for (unsigned e = gpos; e < cap; ++e) lk[e] = KEY_NONE;
      }
      for (unsigned e = c; e < padded; ++e) {
        v.key[off + e] = KEY_NONE;
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) v.d[ir][off + e] = 0.0;
      }
      for (unsigned gq = off >> 3; gq < ((off + padded) >> 3); ++gq) v.grp[gq] = (unsigned char)tid;
    }
    v.run[tid] = run;
  }
  __syncthreads();
}

// Staged groups -> runs at the tails of the global lists: consecutive threads write consecutive
// entries, every group of 8 is one aligned 32 B (keys) / 64 B (residuals) sector.
template <int NR>
__device__ __forceinline__ void part_writeout(const PartView<NR> &v, unsigned cap, unsigned list_first,
                                              unsigned *__restrict__ g_key, double *const (&g_d)[3]) {
  const unsigned total = *v.total;
  for (unsigned s = threadIdx.x * 2u; s < total; s += blockDim.x * 2u) {  // 2 entries per thread: 8 B + 16 B stores
    const unsigned b = v.grp[s >> 3], run = v.run[b];
    if (run == PB_OVERFLOW) continue;
    const size_t gi = (size_t)(list_first + b) * cap + run + (s - v.off[b]);
    *reinterpret_cast<uint2 *>(g_key + gi) = *reinterpret_cast<const uint2 *>(v.key + s);
#pragma unroll
    for (int ir = 0; ir < NR; ++ir)
      *reinterpret_cast<double2 *>(g_d[ir] + gi) = *reinterpret_cast<const double2 *>(v.d[ir] + s);
  }
}

// Direct accumulate of one entry (overflow of a list): the same fp64 RED the direct path uses.
template <int NR>
__device__ __forceinline__ void red_entry(const GridDesc &g, unsigned key, const double (&d)[NR]) {
  double sq[NR];
#pragma unroll
  for (int ir = 0; ir < NR; ++ir) sq[ir] = __dmul_rn(d[ir], d[ir]);
  const double sxy = (NR == 2) ? __dmul_rn(d[0], d[NR - 1]) : 0.0;
  red_record<NR>(g.acc + (size_t)key * (size_t)g.rs, 1.0, d, sq, sxy);
}

// K1p: a batch of P1_BATCH observations per CTA iteration.
template <int NR>
__global__ void __launch_bounds__(P1_THREADS, 2)
k_part_obs(const GridDesc g, const ObsCols c, const FilterDesc f, Status *__restrict__ st, unsigned call_seq,
           const StageDesc sg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PartView<NR> v = part_view<NR, P1_SCAP>(smem_raw);
  double *s_p2 = reinterpret_cast<double *>(smem_raw + ((part_smem_bytes<NR, P1_SCAP>() + 15) & ~(size_t)15));
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid < PB_MAXB) sg.fill1_next[tid] = 0u;  // counters of the NEXT call
  load_p2(g, s_p2);
  const double *p2s = g.p2_in_smem ? s_p2 : g.p2;
  const int nb = (int)sg.nbuckets;
  const int nbatch = (c.n + P1_BATCH - 1) / P1_BATCH;
  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
    if (tid < PB_MAXB) v.cnt[tid] = 0u;
    __syncthreads();
    unsigned rkey[P1_ITEMS], rrank[P1_ITEMS];
    double rd[NR][P1_ITEMS];
#pragma unroll
    for (int h = 0; h < P1_ITEMS / 4; ++h) {
      const int base = batch * P1_BATCH + (h * P1_THREADS + tid) * 4;
      Obs4<NR> o;
      int fa[4];
      if (base < c.n) load_obs4<NR>(c, f, base, o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = h * 4 + j;
        unsigned key = KEY_NONE;
        fa[j] = 0;
        if (base + j < c.n)
          key = classify(g, f, p2s, o.lat[j], o.lon[j], o.lev[j], o.kx[j], o.kt[j], o.fl[j], o.tm[j], base + j, st,
                         call_seq, fa[j]);
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) {
          double d = o.d[ir][j];
          if (f.scale_res && key != KEY_NONE) d = scale_residual(f, c, d, base + j);
          rd[ir][q] = d;
        }
        rkey[q] = key;
        rrank[q] = (key != KEY_NONE) ? atomicAdd(v.cnt + (key >> sg.bucket_shift), 1u) : 0u;
      }
      if (c.flag_out && base < c.n) {
        if (c.aligned && base + 3 < c.n) {
          *reinterpret_cast<int4 *>(c.flag_out + base) = make_int4(fa[0], fa[1], fa[2], fa[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (base + j < c.n) c.flag_out[base + j] = fa[j];
        }
      }
    }
    __syncthreads();
    part_layout<NR>(v, nb, sg.fill1, sg.key1, sg.cap1, 0u);
#pragma unroll
    for (int q = 0; q < P1_ITEMS; ++q) {
      if (rkey[q] == KEY_NONE) continue;
      const unsigned b = rkey[q] >> sg.bucket_shift;
      double d[NR];
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) d[ir] = rd[ir][q];
      if (v.run[b] == PB_OVERFLOW) {
        red_entry<NR>(g, rkey[q], d);
      } else {
        const unsigned s = v.off[b] + rrank[q];
        v.key[s] = rkey[q];
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) v.d[ir][s] = d[ir];
      }
    }
    __syncthreads();
    part_writeout<NR>(v, sg.cap1, 0u, sg.key1, sg.d1);
    __syncthreads();
  }
  finish_status(st, call_seq, c.lev);
}

// K1q: CTA (bucket, chunk) regroups P2_BATCH scratch entries of one bucket by tile.
template <int NR>
__global__ void __launch_bounds__(P2_THREADS, 1)
k_part_tiles(const GridDesc g, const StageDesc sg, int chunks_per_bucket) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PartView<NR> v = part_view<NR, P2_SCAP>(smem_raw);
  const int tid = threadIdx.x;
  const unsigned b = blockIdx.x / (unsigned)chunks_per_bucket, chunk = blockIdx.x % (unsigned)chunks_per_bucket;
  const unsigned n = min(sg.fill1[b], sg.cap1);
  const unsigned start = chunk * (unsigned)P2_BATCH;
  if (start >= n) return;
  const int tsh = sg.bucket_shift - sg.tile_shift;
  const int nt = 1 << tsh;                        // tiles per bucket (<= PB_MAXB)
  if (tid < PB_MAXB) v.cnt[tid] = 0u;
  __syncthreads();
  const size_t src = (size_t)b * sg.cap1 + start;
  unsigned rkey[P2_ITEMS], rrank[P2_ITEMS];
  double rd[NR][P2_ITEMS];
#pragma unroll
  for (int q = 0; q < P2_ITEMS; ++q) {
    const unsigned e = (unsigned)(q * P2_THREADS + tid);
    rkey[q] = (start + e < n) ? sg.key1[src + e] : KEY_NONE;
  }
#pragma unroll
  for (int q = 0; q < P2_ITEMS; ++q) {
    const unsigned e = (unsigned)(q * P2_THREADS + tid);
    const bool live = rkey[q] != KEY_NONE;
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) rd[ir][q] = live ? sg.d1[ir][src + e] : 0.0;
    rrank[q] = live ? atomicAdd(v.cnt + ((rkey[q] >> sg.tile_shift) & (unsigned)(nt - 1)), 1u) : 0u;
  }
  __syncthreads();
  part_layout<NR>(v, nt, sg.fill, sg.skey, sg.cap, b << tsh);
#pragma unroll
  for (int q = 0; q < P2_ITEMS; ++q) {
    if (rkey[q] == KEY_NONE) continue;
    const unsigned t = (rkey[q] >> sg.tile_shift) & (unsigned)(nt - 1);
    double d[NR];
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) d[ir] = rd[ir][q];
    if (v.run[t] == PB_OVERFLOW) {
      red_entry<NR>(g, rkey[q], d);
    } else {
      const unsigned s = v.off[t] + rrank[q];
      v.key[s] = rkey[q];
#pragma unroll
      for (int ir = 0; ir < NR; ++ir) v.d[ir][s] = d[ir];
    }
  }
  __syncthreads();
  part_writeout<NR>(v, sg.cap, b << tsh, sg.skey, sg.sd);
}

// ---------------------------------------------------------------------------------------------
// Tiled fast path, part 2 (K2t): one CTA per tile.  The tile's count / sum / sum-of-squares planes
// live in shared memory (privatised histogram); the tile's list streams in coalesced; lanes of a warp
// that hit the same box are combined first (__match_any_sync) and only then touch shared memory;
// the tile is finally added to the HBM records with plain coalesced read-modify-writes (the tile is
// owned by this CTA: no global atomics).
template <int NR>
struct TilePlanes { static constexpr int NV = (NR == 1) ? 2 : (NR == 2 ? 5 : 6); };
constexpr int TILE_THREADS = 512;
constexpr int TILE_ILP = 2;

template <int NR>
__global__ void __launch_bounds__(TILE_THREADS, 2)
k_tile_accum(const GridDesc g, const StageDesc sg, unsigned nrec) {
  constexpr int NV = TilePlanes<NR>::NV;
  extern __shared__ double s_val[];                       // NV planes of T doubles, then T counts
  const unsigned T = 1u << sg.tile_shift;
  unsigned *s_cnt = reinterpret_cast<unsigned *>(s_val + (size_t)NV * T);
  const int rs_shift = (g.rs == 4) ? 2 : 3;
  for (unsigned tile = blockIdx.x; tile < sg.ntiles; tile += gridDim.x) {
    const unsigned n = min(sg.fill[tile], sg.cap);
    if (n == 0) continue;                                 // uniform across the CTA
    for (unsigned e = threadIdx.x; e < NV * T; e += blockDim.x) s_val[e] = 0.0;
    for (unsigned e = threadIdx.x; e < T; e += blockDim.x) s_cnt[e] = 0u;
    __syncthreads();
    const size_t base = (size_t)tile * sg.cap;
    const unsigned stride = blockDim.x * TILE_ILP;
    const unsigned nround = (n + stride - 1) / stride * stride;   // whole warps iterate together
    for (unsigned r0 = threadIdx.x; r0 < nround; r0 += stride) {
      unsigned bx[TILE_ILP];
      double d[TILE_ILP][NR];
#pragma unroll
      for (int u = 0; u < TILE_ILP; ++u) {                // all loads of the round first
        const unsigned r = r0 + u * blockDim.x;
        const unsigned key = (r < n) ? __ldcs(sg.skey + base + r) : KEY_NONE;
        bx[u] = (key == KEY_NONE) ? 0xFFFFFFFFu : (key & (T - 1u));
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) d[u][ir] = (r < n) ? __ldcs(sg.sd[ir] + base + r) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < TILE_ILP; ++u) {
        const unsigned b = bx[u];
        const bool live = b != 0xFFFFFFFFu;
        double v[NV];
        if (NR == 1) { v[0] = d[u][0]; v[1] = __dmul_rn(d[u][0], d[u][0]); }
        if (NR == 2) {
          v[0] = d[u][0]; v[1] = __dmul_rn(d[u][0], d[u][0]);
          v[2] = d[u][NR - 1]; v[3] = __dmul_rn(d[u][NR - 1], d[u][NR - 1]);
          v[4] = __dmul_rn(d[u][0], d[u][NR - 1]);
        }
        if (NR == 3) {
#pragma unroll
          for (int ir = 0; ir < NR; ++ir) { v[2 * ir] = d[u][ir]; v[2 * ir + 1] = __dmul_rn(d[u][ir], d[u][ir]); }
        }
        unsigned cnt = 1u;
        bool issue = live;
        // warp aggregation: only pays (and only runs) when lanes collide, i.e. clustered data
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        const bool multi = live && (peers & (peers - 1u));
        if (__any_sync(0xffffffffu, multi)) {
          const int lane = threadIdx.x & 31;
          const int leader = __ffs(peers) - 1;
          const int npeer = live ? __popc(peers) : 1;
          const int maxc = __reduce_max_sync(0xffffffffu, npeer);
          unsigned others = peers & ~(1u << leader);
          for (int it = 1; it < maxc; ++it) {
            const int src = others ? (__ffs(others) - 1) : lane;
            others &= others - 1u;
            const bool take = (lane == leader) && (it < npeer);
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              const double a = __shfl_sync(0xffffffffu, v[q], src);
              if (take) v[q] = __dadd_rn(v[q], a);
            }
          }
          cnt = (unsigned)npeer;
          issue = live && (lane == leader);
        }
        if (issue) {
          atomicAdd(s_cnt + b, cnt);
#pragma unroll
          for (int q = 0; q < NV; ++q) atomicAdd(s_val + (size_t)q * T + b, v[q]);
        }
      }
    }
    __syncthreads();
    // add the tile to the HBM records: element e of the tile's span is (box e >> rs_shift, slot e & (rs-1));
    // each thread handles pairs of doubles (16 B), four pairs in flight
    const size_t first = (size_t)tile << sg.tile_shift;
    const unsigned nb = (unsigned)min((size_t)T, (size_t)nrec - first);
    double2 *span = reinterpret_cast<double2 *>(g.acc + first * (size_t)g.rs);
    const unsigned npair = (nb << rs_shift) >> 1;
    for (unsigned e0 = threadIdx.x; e0 < npair; e0 += blockDim.x * 4u) {
      double2 cur[4], add[4];
      bool any[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned e = e0 + u * blockDim.x;
        any[u] = false;
        if (e < npair) {
          const unsigned el = e << 1, bxi = el >> rs_shift, slot = el & ((1u << rs_shift) - 1u);
          const unsigned cn = s_cnt[bxi];
          if (cn != 0u && slot <= (unsigned)NV) {
            any[u] = true;
            add[u].x = (slot == 0u) ? (double)cn : s_val[(size_t)(slot - 1u) * T + bxi];
            add[u].y = (slot + 1u <= (unsigned)NV) ? s_val[(size_t)slot * T + bxi] : 0.0;
            cur[u] = span[e];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (any[u]) {
          cur[u].x = __dadd_rn(cur[u].x, add[u].x);
          cur[u].y = __dadd_rn(cur[u].y, add[u].y);
          span[e0 + u * blockDim.x] = cur[u];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) sg.fill[tile] = 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// K3 finalize.  Output planes per box, in this order: bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs.
struct OutPlanes {
  double *bias, *rtms, *stdv, *xcov, *nobs;  // device arrays in the reference's layout (may be null)
  int xcov_planes;                           // how many ir slabs of xcov exist (1 for the raw getters)
};

// m_gritas_grids.f:566-605 (2-D) and 612-654 (3-D) for one box.  raw: pass the sums through.
template <int NR>
__device__ __forceinline__ void finalize_box(const double *__restrict__ rec, bool is3d, int nrmlz, int ntresh,
                                             bool raw, double (&bias)[NR], double (&rtms)[NR], double (&stdv)[NR],
                                             double (&xcov)[NR], double &nobs) {
  const double n = rec[0];
  nobs = n;
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
  double tmp[NR], xtmp = 0.0;
  if (some) {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) {
      bias[ir] = __ddiv_rn(rec[1 + 2 * ir], dm);
      const double t = __ddiv_rn(rec[2 + 2 * ir], dm);
      rtms[ir] = __dsqrt_rn(t);
      tmp[ir] = fabs(__dsub_rn(t, __dmul_rn(bias[ir], bias[ir])));
    }
    xtmp = __ddiv_rn(xraw, dm);
    xcov[0] = xtmp;
    xtmp = __dsub_rn(xtmp, __dmul_rn(bias[0], bias[nx - 1]));
  } else {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) { bias[ir] = AMISS; rtms[ir] = AMISS; tmp[ir] = 0.0; }
    xcov[0] = AMISS;
  }
  if (many) {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) stdv[ir] = __dsqrt_rn(__ddiv_rn(__dmul_rn(dm, tmp[ir]), dm1));
    xcov[nx - 1] = __ddiv_rn(__dmul_rn(dm, xtmp), dm1);
  } else if (is3d && n == 1.0) {  // m_gritas_grids.f:642-644 (the 2-D branch has no such case)
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) stdv[ir] = __dsqrt_rn(tmp[ir]);
    xcov[nx - 1] = xtmp;
  } else {
#pragma unroll
    for (int ir = 0; ir < NR; ++ir) stdv[ir] = AMISS;
    xcov[nx - 1] = AMISS;
  }
}

// 3-D: tile of 32 levels x 32 longitudes for one (j,l).  Records are read level-fastest (the HBM
// layout), statistics computed in registers, transposed through shared memory one plane at a
// time, written longitude-fastest (the reference's layout).
template <int NR>
__global__ void __launch_bounds__(1024)
k_finalize3(const GridDesc g, OutPlanes out, int nrmlz, int ntresh, int raw) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  const size_t plane = (size_t)g.im * g.jm * g.km * g.l3d;  // one ir slab
  for (int jl = blockIdx.z; jl < g.jm * g.l3d; jl += gridDim.z) {
    const int j = jl % g.jm, l = jl / g.jm;
    double v[4 * NR + 1];
#pragma unroll
    for (int p = 0; p < 4 * NR + 1; ++p) v[p] = 0.0;
    {
      const int k = k0 + tx, i = i0 + ty;
      if (k < g.nlevs && i < g.im) {  // levels nlevs..km-1 are never touched by the reference: zeros
        const double *rec =
            g.acc + ((size_t)k + (size_t)g.km * ((size_t)i + (size_t)g.im * ((size_t)j + (size_t)g.jm * l))) * g.rs;
        double r[8];
        const double2 a = *reinterpret_cast<const double2 *>(rec), b = *reinterpret_cast<const double2 *>(rec + 2);
        r[0] = a.x; r[1] = a.y; r[2] = b.x; r[3] = b.y;
        if (NR > 1) {
          const double2 c2 = *reinterpret_cast<const double2 *>(rec + 4), d2 = *reinterpret_cast<const double2 *>(rec + 6);
          r[4] = c2.x; r[5] = c2.y; r[6] = d2.x; r[7] = d2.y;
        }
        double bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs;
        finalize_box<NR>(r, true, nrmlz, ntresh, raw != 0, bias, rtms, stdv, xcov, nobs);
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) {
          v[ir] = bias[ir]; v[NR + ir] = rtms[ir]; v[2 * NR + ir] = stdv[ir]; v[3 * NR + ir] = xcov[ir];
        }
        v[4 * NR] = nobs;
      }
    }
    const int ko = k0 + ty, io = i0 + tx;
    const bool inb = ko < g.km && io < g.im;
    const size_t o = (size_t)io + (size_t)g.im * ((size_t)j + (size_t)g.jm * ((size_t)ko + (size_t)g.km * l));
#pragma unroll
    for (int p = 0; p < 4 * NR + 1; ++p) {
      double *dst = nullptr;
      const int grp = p / NR, ir = p % NR;
      if (p == 4 * NR) dst = out.nobs;
      else if (grp == 0) dst = out.bias;
      else if (grp == 1) dst = out.rtms;
      else if (grp == 2) dst = out.stdv;
      else dst = out.xcov;
      if (dst == nullptr) continue;  // uniform across the block
      if (grp == 3 && ir >= out.xcov_planes) continue;
      tile[ty][tx] = v[p];
      __syncthreads();
      if (inb) dst[o + (p == 4 * NR ? 0 : plane * ir)] = tile[tx][ty];
      __syncthreads();
    }
  }
}

// 2-D: one thread per surface box, no transpose needed.
template <int NR>
__global__ void __launch_bounds__(256)
k_finalize2(const GridDesc g, OutPlanes out, int nrmlz, int ntresh, int raw) {
  const size_t nb = g.nbox2;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (size_t)gridDim.x * blockDim.x) {
    const double *rec = g.acc + ((size_t)g.nbox3 + b) * g.rs;
    double r[8];
#pragma unroll
    for (int q = 0; q < (NR > 1 ? 8 : 4); ++q) r[q] = rec[q];
    double bias[NR], rtms[NR], stdv[NR], xcov[NR], nobs;
    finalize_box<NR>(r, false, nrmlz, ntresh, raw != 0, bias, rtms, stdv, xcov, nobs);
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

}  // namespace gb200
