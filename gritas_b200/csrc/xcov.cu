// xcov.cu -- channel cross-covariance family of m_gritas_grids on the GPU.
//
//   gritas_b200_xcov_accum   GrITAS_XCov_Accum_   m_gritas_grids.f:667-848
//   gritas_b200_xcov_norm    GrITAS_XCov_Norm_    m_gritas_grids.f:859-1008
//
// The reference walks the observations profile by profile (km consecutive observations), keeps the profiles whose
// active channels are all present and accepted, and adds the outer product of the profile's residual vectors to
// xcov_lv(nchan, nchan, l3d, nr) -- a rank-1 update per profile, i.e. X^T Y over the accepted profiles.  Here:
//   k_xc_scan     one warp per profile: channel match, ob_flag / table test, accept flag (ob_flag write-back as :795-798)
//   k_xc_rank     exclusive scan of the accept flags: accepted profiles keep their order
//   k_xc_gather   one warp per accepted profile: dense rows X[ir][q][slot] = del(indx(slot), ir) and the grid index l
//   k_xc_outer    64 x 64 output tiles, 4 x 4 per thread, the accepted profiles streamed through shared memory IN ORDER;
//                 every element is updated with one multiply and one add per profile (no FMA, no re-association):
//                 the same sequence of fp64 operations as the reference loop, so the sums are bit-identical to it
//   k_xc_bias     bias_lv / nobs_lv, one thread per (slot, l, ir), profiles in order
//   k_xc_norm*    the Norm, element-wise on copies of the accumulators (the device state is left untouched)
// fp64 vector pipe (no tensor cores: tcgen05 has no fp64 kind, and DMMA would re-associate the sums).
#include "../../include/gritas_b200.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>
#include <cub/device/device_radix_sort.cuh>

namespace {

struct XcDesc {
  int km, nchan, l3d, nr, tch;
  int kxmax, ktmax;
  int lut_n;                    // channel numbers 0 .. lut_n-1 have an entry in lut
  const unsigned char *lut;     // 1: the channel is active (is in achan)
  const int *table;
};

// Fortran NINT of a 64-bit real into a default integer, as oracle/ and kernels.cuh pin it
__device__ __forceinline__ long long xc_nint(double x) {
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

// One warp per profile (m_gritas_grids.f:774-802).  accept[p] = 1 iff every active channel of the profile is present
// exactly once-per-slot (m1 == nchan) and none of them is rejected.  too_many: a profile matched more than nchan
// observations (:789-790, fatal in the reference).
__global__ void __launch_bounds__(256)
k_xc_scan(const XcDesc d, int nprof, const double *__restrict__ lev, const int *__restrict__ kx, const int *__restrict__ kt,
          int *__restrict__ ob_flag, unsigned *__restrict__ accept, int *__restrict__ too_many) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < nprof; p += gridDim.x * wpb) {
    const size_t base = (size_t)p * d.km;
    int m1 = 0;
    bool rej = false;
    for (int o = 0; o < d.km; o += 32) {
      const int ii = o + lane;
      bool act = false;
      if (ii < d.km) {
        const long long c = xc_nint(lev[base + ii]);
        act = c >= 0 && c < d.lut_n && d.lut[c] != 0;                    // :780-786
      }
      const unsigned bal = __ballot_sync(0xffffffffu, act);
      if (act) {
        const int slot = m1 + __popc(bal & ((1u << lane) - 1u));
        if (slot < d.nchan) {                                            // beyond nchan: fatal, see below
          bool r = ob_flag && ob_flag[base + ii] != 0;                   // :793
          const int x = kx[base + ii], t = kt[base + ii];
          int l = 0;
          if (x >= 1 && x <= d.kxmax && t >= 1 && t <= d.ktmax) l = d.table[(size_t)(x - 1) + (size_t)d.kxmax * (size_t)(t - 1)];
          l = abs(l);                                                    // :794-795
          if (l <= 0) {                                                  // :796-799
            if (ob_flag) ob_flag[base + ii] = 1;
            r = true;
          }
          rej |= r;
        }
      }
      m1 += __popc(bal);
    }
    rej = __any_sync(0xffffffffu, rej);
    if (lane == 0) {
      if (m1 > d.nchan) atomicExch(too_many, 1);
      accept[p] = (!rej && m1 == d.nchan) ? 1u : 0u;                     // :801-802
    }
  }
}

// exclusive scan of accept[0 .. n) into rank[], total into *count (one block; n is a few 10^4)
__global__ void __launch_bounds__(1024)
k_xc_rank(const unsigned *__restrict__ accept, unsigned *__restrict__ rank, int n, int *__restrict__ count) {
  __shared__ unsigned s_w[32];
  __shared__ unsigned s_base;
  if (threadIdx.x == 0) s_base = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 0; o < n; o += 1024) {
    const int i = o + threadIdx.x;
    const unsigned v = i < n ? accept[i] : 0u;
    unsigned inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, s);
      if (lane >= s) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const unsigned w = s_w[lane];
      unsigned wi = w;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, s);
        if (lane >= s) wi += t;
      }
      s_w[lane] = wi - w;
    }
    __syncthreads();
    const unsigned excl = s_base + s_w[wid] + inc - v;
    if (i < n) rank[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) s_base = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = (int)s_base;
}

// One warp per accepted profile: rows of the dense operand.  X[ir][q][slot] (slot fastest), L[q][slot] = grid index (0-based).
__global__ void __launch_bounds__(256)
k_xc_gather(const XcDesc d, int nprof, int nobs, const double *__restrict__ lev, const int *__restrict__ kx,
            const int *__restrict__ kt, const double *__restrict__ del, const unsigned *__restrict__ accept,
            const unsigned *__restrict__ rank, int nacc, double *__restrict__ X, int *__restrict__ L) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < nprof; p += gridDim.x * wpb) {
    if (!accept[p]) continue;
    const size_t q = rank[p];
    const size_t base = (size_t)p * d.km;
    int m1 = 0;
    for (int o = 0; o < d.km; o += 32) {
      const int ii = o + lane;
      bool act = false;
      if (ii < d.km) {
        const long long c = xc_nint(lev[base + ii]);
        act = c >= 0 && c < d.lut_n && d.lut[c] != 0;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, act);
      if (act) {
        const int slot = m1 + __popc(bal & ((1u << lane) - 1u));       // indx(m1) = ii (:791)
        const int l = abs(d.table[(size_t)(kx[base + ii] - 1) + (size_t)d.kxmax * (size_t)(kt[base + ii] - 1)]);
        L[q * d.nchan + slot] = l - 1;
        for (int ir = 0; ir < d.nr; ++ir)                                // rdel(:,:) = del(indx(:),:)  (:809)
          X[((size_t)ir * nacc + q) * d.nchan + slot] = del[base + ii + (size_t)nobs * ir];
      }
      m1 += __popc(bal);
    }
  }
}

// xcov_lv(ii, kk, l, s) += A[p][ii] * B[p][kk] for the accepted profiles p in order, where l = L[p][kk]
// (m_gritas_grids.f:815-831).  Block = 64 x 64 outputs of one (l, s); thread = 4 x 4.
constexpr int XT = 64, XP = 16;
__global__ void __launch_bounds__(256)
k_xc_outer(const XcDesc d, int nacc, const double *__restrict__ X, const int *__restrict__ L, double *__restrict__ xcov) {
  __shared__ double s_a[XP][XT], s_b[XP][XT];
  __shared__ int s_l[XP][XT];
  const int l = blockIdx.z % d.l3d, s = blockIdx.z / d.l3d;
  const int sa = d.tch ? s : 0, sb = d.tch ? s : 1;                       // tch: rdel(ii,:)*rdel(kk,:); else rdel(ii,1)*rdel(kk,2)
  const int i0 = blockIdx.x * XT, k0 = blockIdx.y * XT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;                 // ii = i0 + tx + 16*u, kk = k0 + ty + 16*v
  double acc[4][4];
  const size_t plane = (size_t)d.nchan * d.nchan;
  double *out = xcov + plane * ((size_t)l + (size_t)d.l3d * s);
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int ii = i0 + tx + 16 * u, kk = k0 + ty + 16 * v;
      acc[u][v] = (ii < d.nchan && kk < d.nchan) ? out[(size_t)ii + (size_t)d.nchan * kk] : 0.0;
    }
  const double *A = X + (size_t)sa * nacc * d.nchan, *B = X + (size_t)sb * nacc * d.nchan;
  for (int p0 = 0; p0 < nacc; p0 += XP) {
    for (int e = threadIdx.x; e < XP * XT; e += 256) {
      const int pp = e / XT, c = e % XT;
      const bool inp = p0 + pp < nacc;
      s_a[pp][c] = (inp && i0 + c < d.nchan) ? A[(size_t)(p0 + pp) * d.nchan + i0 + c] : 0.0;
      s_b[pp][c] = (inp && k0 + c < d.nchan) ? B[(size_t)(p0 + pp) * d.nchan + k0 + c] : 0.0;
      s_l[pp][c] = (inp && k0 + c < d.nchan) ? L[(size_t)(p0 + pp) * d.nchan + k0 + c] : -1;
    }
    __syncthreads();
    const int np = min(XP, nacc - p0);
    for (int pp = 0; pp < np; ++pp) {
      double a[4], b[4];
      bool on[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = s_a[pp][tx + 16 * u];
#pragma unroll
      for (int v = 0; v < 4; ++v) { b[v] = s_b[pp][ty + 16 * v]; on[v] = s_l[pp][ty + 16 * v] == l; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (on[v]) acc[u][v] = __dadd_rn(acc[u][v], __dmul_rn(a[u], b[v]));
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int ii = i0 + tx + 16 * u, kk = k0 + ty + 16 * v;
      if (ii < d.nchan && kk < d.nchan) out[(size_t)ii + (size_t)d.nchan * kk] = acc[u][v];
    }
}

// bias_lv(kk, l, ir) += rdel(kk, ir), nobs_lv(kk, l) += 1 for the accepted profiles in order (:817, :832)
__global__ void __launch_bounds__(256)
k_xc_bias(const XcDesc d, int nacc, const double *__restrict__ X, const int *__restrict__ L, double *__restrict__ bias,
          double *__restrict__ nobs) {
  const int total = d.nchan * d.l3d * (d.nr + 1);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kk = e % d.nchan, l = (e / d.nchan) % d.l3d, ir = e / (d.nchan * d.l3d);   // ir == nr: the counts
    double *dst = ir < d.nr ? bias + (size_t)kk + (size_t)d.nchan * ((size_t)l + (size_t)d.l3d * ir) : nobs + (size_t)kk + (size_t)d.nchan * l;
    double v = *dst;
    for (int q = 0; q < nacc; ++q) {
      if (L[(size_t)q * d.nchan + kk] != l) continue;
      v = __dadd_rn(v, ir < d.nr ? X[((size_t)ir * nacc + q) * d.nchan + kk] : 1.0);
    }
    *dst = v;
  }
}

// GrITAS_XCov_Norm_ part 1 (:938-958): the means.  m is an INTEGER (:903), `m = n` truncates.
__global__ void __launch_bounds__(256)
k_xc_norm_bias(const XcDesc d, int nrmlz, int ntresh, const double *__restrict__ nobs, double *__restrict__ bias) {
  const int total = d.nchan * d.l3d * d.nr;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int k2 = e % d.nchan, l = (e / d.nchan) % d.l3d;
    const int m = nrmlz ? (int)nobs[(size_t)k2 + (size_t)d.nchan * l] : 1;
    if (m > 1 && m >= ntresh) bias[e] = __ddiv_rn(bias[e], (double)m);
  }
}

// part 2 (:959-989): covariances from the normalised means; part 3 (:994-1002): the three-cornered hat.
__global__ void __launch_bounds__(256)
k_xc_norm_cov(const XcDesc d, int nrmlz, int ntresh, int self, const double *__restrict__ nobs, const double *__restrict__ bias,
              double *__restrict__ xcov) {
  const size_t plane = (size_t)d.nchan * d.nchan * d.l3d, nb = (size_t)d.nchan * d.l3d;
  const int nrr = d.tch ? 3 : 1;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x) {
    const int k1 = (int)(e % d.nchan), k2 = (int)((e / d.nchan) % d.nchan), l = (int)(e / ((size_t)d.nchan * d.nchan));
    const int m = nrmlz ? (int)nobs[(size_t)k2 + (size_t)d.nchan * l] : 1;
    double v[3] = {0.0, 0.0, 0.0};
    for (int k3 = 0; k3 < nrr; ++k3) {
      double x = xcov[e + plane * k3];
      if (m > 1 && m >= ntresh) {
        const int i1 = d.tch ? k3 : 0, i2 = d.tch ? k3 : 1;
        const double b1a = bias[(size_t)k1 + (size_t)d.nchan * l + nb * i1], b2b = bias[(size_t)k2 + (size_t)d.nchan * l + nb * i2];
        double xt = __ddiv_rn(x, (double)m);
        if (d.tch) {
          xt = __dsub_rn(xt, __dmul_rn(b1a, b2b));
        } else {
          const double b2a = bias[(size_t)k2 + (size_t)d.nchan * l + nb * i1], b1b = bias[(size_t)k1 + (size_t)d.nchan * l + nb * i2];
          xt = __dsub_rn(xt, __dmul_rn(0.5, __dadd_rn(__dmul_rn(b1a, b2b), __dmul_rn(b2a, b1b))));
        }
        x = __ddiv_rn(__dmul_rn((double)m, xt), (double)(m - 1));
      }
      v[k3] = x;
    }
    if (d.tch && !self) {
      xcov[e] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v[0], v[1]), v[2]));
      xcov[e + plane] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v[2], v[0]), v[1]));
      xcov[e + 2 * plane] = __dmul_rn(0.5, __dsub_rn(__dadd_rn(v[1], v[2]), v[0]));
    } else {
      for (int k3 = 0; k3 < nrr; ++k3) xcov[e + plane * k3] = v[k3];
    }
  }
}


// ---- GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059): level cross-covariances of conventional soundings ---------------
// Profiles are runs of equal sounding index ks; every run gathers ALL observations that carry its index (:1973-1979), so an
// index that comes back later is processed once per run -- as in the reference.  ndim = nlevs, grid l = 1.

// level of every observation (:2004-2013, literal scan of the GrITAS_Pint intervals); the earliest miss is recorded
__global__ void __launch_bounds__(256)
k_xp_levels(int n, int nlevs, const double *__restrict__ lev, const double *__restrict__ p1, const double *__restrict__ p2,
            int *__restrict__ kidx, int *__restrict__ first_miss) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = lev[i];
    int k = -1;
    for (int kk = 0; kk < nlevs; ++kk)
      if (x <= p1[kk] && x > p2[kk]) { k = kk; break; }
    kidx[i] = k;
    if (k < 0) atomicMin(first_miss, i);
  }
}

// run starts (:1942-1964) and the order-preserving sort key of ks
__global__ void __launch_bounds__(256)
k_xp_runs(int n, const int *__restrict__ ks, unsigned *__restrict__ start, unsigned *__restrict__ key, int *__restrict__ idx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    start[i] = (i == 0 || ks[i] != ks[i - 1]) ? 1u : 0u;
    key[i] = (unsigned)ks[i] ^ 0x80000000u;
    idx[i] = i;
  }
}

// compaction of the run starts: run_ks[r] = ks at the r-th start
__global__ void __launch_bounds__(256)
k_xp_run_ks(int n, const int *__restrict__ ks, const unsigned *__restrict__ start, const unsigned *__restrict__ rank,
            unsigned *__restrict__ run_key) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (start[i]) run_key[rank[i]] = (unsigned)ks[i] ^ 0x80000000u;
}

// GROUP = all observations of one sounding index, in input order after the stable sort; groups are the distinct keys of the sorted list.
__global__ void __launch_bounds__(256)
k_xp_gstart(int n, const unsigned *__restrict__ skey, unsigned *__restrict__ gstart) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) gstart[i] = (i == 0 || skey[i] != skey[i - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
k_xp_groups(int n, const unsigned *__restrict__ gstart, const unsigned *__restrict__ grank, int ngroups, unsigned *__restrict__ gbeg) {
  // group begin positions in the sorted list
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (gstart[i]) gbeg[grank[i]] = (unsigned)i;
  if (blockIdx.x == 0 && threadIdx.x == 0) gbeg[ngroups] = (unsigned)n;
}

// one thread per group: its members bucketed by level, stably, into lm[]; goff[g][k] = start of level k (positions in lm)
__global__ void __launch_bounds__(256)
k_xp_bucket(int nlevs, const int *__restrict__ sidx, const int *__restrict__ kidx, int ngroups, const unsigned *__restrict__ gbeg,
            unsigned *__restrict__ goff, int *__restrict__ lm) {
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
    const unsigned a = gbeg[g], b = gbeg[g + 1];
    unsigned pos = a;
    for (int k = 0; k < nlevs; ++k) {
      goff[(size_t)g * (nlevs + 1) + k] = pos;
      for (unsigned q = a; q < b; ++q) {
        const int i = sidx[q];
        if (kidx[i] == k) lm[pos++] = i;
      }
    }
    goff[(size_t)g * (nlevs + 1) + nlevs] = pos;
  }
}

// run -> group: position of the run's key among the distinct sorted keys
__global__ void __launch_bounds__(256)
k_xp_run_group(int nruns, int ngroups, const unsigned *__restrict__ run_key, const unsigned *__restrict__ skey,
               const unsigned *__restrict__ gbeg, int *__restrict__ run_group) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nruns; r += gridDim.x * blockDim.x) {
    const unsigned key = run_key[r];
    int lo = 0, hi = ngroups - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (skey[gbeg[mid]] < key) lo = mid + 1; else hi = mid;
    }
    run_group[r] = lo;
  }
}

// One thread per cell (k1, k2, series): the runs in order, the level-k1 members of the run in order, the level-k2 members in
// order -- the (n, m) order of the reference's nested loops (:2027-2049) restricted to this cell: bit-identical sums.
// Cells with k2 == 0 also carry bias_lv(k1,1,:) and nobs_lv(k1,1) (:2034-2035).
__global__ void __launch_bounds__(128)
k_xp_accum(int nlevs, int l3d, int nr, int tch, int nobs, int nruns, const int *__restrict__ run_group, const unsigned *__restrict__ goff,
           const int *__restrict__ lm, const double *__restrict__ del, double *__restrict__ bias, double *__restrict__ xcov,
           double *__restrict__ nobsv) {
  const int nser = tch ? nr : 1;
  const int total = nlevs * nlevs * nser;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int k2 = e % nlevs, k1 = (e / nlevs) % nlevs, s = e / (nlevs * nlevs);
    double *xc = xcov + (size_t)k1 + (size_t)nlevs * ((size_t)k2 + (size_t)nlevs * ((size_t)0 + (size_t)l3d * s));
    double acc = *xc;
    const bool edge = k2 == 0;                                   // this thread also owns bias(k1, 1, s...) and nobs(k1, 1)
    double bsum[3] = {0.0, 0.0, 0.0}, cnt = 0.0;
    if (edge && s == 0) {
      for (int ir = 0; ir < nr; ++ir) bsum[ir] = bias[(size_t)k1 + (size_t)nlevs * ((size_t)0 + (size_t)l3d * ir)];
      cnt = nobsv[k1];
    }
    for (int r = 0; r < nruns; ++r) {
      const unsigned *go = goff + (size_t)run_group[r] * (nlevs + 1);
      const unsigned a0 = go[k1], a1 = go[k1 + 1];
      if (a1 == a0) continue;
      const unsigned b0 = go[k2], b1 = go[k2 + 1];
      for (unsigned a = a0; a < a1; ++a) {
        const int n = lm[a];
        if (edge && s == 0) {
          cnt = __dadd_rn(cnt, 1.0);
          for (int ir = 0; ir < nr; ++ir) bsum[ir] = __dadd_rn(bsum[ir], del[(size_t)n + (size_t)nobs * ir]);
        }
        for (unsigned b = b0; b < b1; ++b) {
          const int m = lm[b];
          if (tch) {
            acc = __dadd_rn(acc, __dmul_rn(del[(size_t)n + (size_t)nobs * s], del[(size_t)m + (size_t)nobs * s]));
          } else {
            const double t1 = __dmul_rn(del[n], del[(size_t)m + nobs]), t2 = __dmul_rn(del[(size_t)n + nobs], del[m]);
            acc = __dadd_rn(acc, __dmul_rn(0.5, __dadd_rn(t1, t2)));
          }
        }
      }
    }
    *xc = acc;
    if (edge && s == 0) {
      for (int ir = 0; ir < nr; ++ir) bias[(size_t)k1 + (size_t)nlevs * ((size_t)0 + (size_t)l3d * ir)] = bsum[ir];
      nobsv[k1] = cnt;
    }
  }
}

}  // namespace

struct gritas_b200_xcov_ctx {
  int device = 0;
  XcDesc d = {};
  cudaStream_t st = nullptr;
  unsigned char *d_lut = nullptr;
  int *d_table = nullptr;
  double *bias = nullptr, *xcov = nullptr, *nobs = nullptr;     // accumulators (nchan,l3d,nr) (nchan,nchan,l3d,nr) (nchan,l3d)
  double *w_bias = nullptr, *w_xcov = nullptr;                   // Norm works on copies
  size_t n_bias = 0, n_xcov = 0, n_nobs = 0;
  // per-call scratch
  void *cols = nullptr; size_t cols_cap = 0;
  unsigned *accept = nullptr, *rank = nullptr; size_t prof_cap = 0, rank_cap = 0;
  double *X = nullptr; size_t x_cap = 0;
  int *L = nullptr; size_t l_cap = 0;
  int *d_small = nullptr;                                         // [0] accepted count, [1] too-many flag
  // pressure-level variant (GrITAS_XCov_Prs_Accum_): the handle was made by gritas_b200_xcov_prs_create
  bool prs = false;
  int nlevs = 0;
  double *d_p1 = nullptr, *d_p2 = nullptr;
  void *prs_scr = nullptr; size_t prs_cap = 0;
  void *prs_tmp = nullptr; size_t prs_tmp_cap = 0;
  uint64_t launches = 0;
  char err[256] = {0};
};

namespace {
int xfail(gritas_b200_xcov_handle h, int rc, const char *fmt, ...) {
  if (h) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(h->err, sizeof(h->err), fmt, ap);
    va_end(ap);
  }
  return rc;
}
#define XCU(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) return xfail(h, GRITAS_B200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

template <class T>
int xgrow(gritas_b200_xcov_handle h, T **p, size_t *cap, size_t count) {
  if (*cap >= count) return 0;
  if (*p) XCU(cudaFree(*p));
  *p = nullptr;
  *cap = 0;
  const size_t want = count + count / 8 + 64;
  XCU(cudaMalloc((void **)p, want * sizeof(T)));
  *cap = want;
  return 0;
}

bool is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
}  // namespace

extern "C" {

int gritas_b200_xcov_create(gritas_b200_xcov_handle *hp, int km, int nchan, const int32_t *achan, int l3d, int nr, int tch,
                            const int32_t *kxkt_table, int kxmax, int ktmax, int device) {
  if (!hp) return GRITAS_B200_ERR_ARG;
  *hp = nullptr;
  if (km < 1 || nchan < 1 || nchan > km || !achan || l3d < 1 || !kxkt_table || kxmax < 1 || ktmax < 1) return GRITAS_B200_ERR_ARG;
  if (nr > 3 || nr < 2) return GRITAS_B200_ERR_NR;                  // 'improper dim settings' (:734-736)
  if (tch ? nr != 3 : nr != 2) return GRITAS_B200_ERR_NR;           // ':TCH / Desroziers setting inconsistent w/ dims' (:737-745)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return GRITAS_B200_ERR_CUDA; }   // no CPU fallback
  if (device < 0) {
    const char *lr = getenv("LOCAL_RANK");
    if (lr && *lr) device = atoi(lr) % ndev;
    else if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return GRITAS_B200_ERR_ARG;
  gritas_b200_xcov_handle h = new (std::nothrow) gritas_b200_xcov_ctx();
  if (!h) return GRITAS_B200_ERR_ARG;
  h->device = device;
  XcDesc &d = h->d;
  d.km = km; d.nchan = nchan; d.l3d = l3d; d.nr = nr; d.tch = tch ? 1 : 0; d.kxmax = kxmax; d.ktmax = ktmax;
  int cmax = 0;
  for (int k = 0; k < nchan; ++k) {
    if (achan[k] < 0 || achan[k] > (1 << 24)) { delete h; return GRITAS_B200_ERR_ARG; }
    cmax = std::max(cmax, (int)achan[k]);
  }
  std::vector<unsigned char> lut((size_t)cmax + 1, 0);
  for (int k = 0; k < nchan; ++k) lut[achan[k]] = 1;
  d.lut_n = cmax + 1;
  h->n_bias = (size_t)nchan * l3d * nr;
  h->n_xcov = (size_t)nchan * nchan * l3d * nr;
  h->n_nobs = (size_t)nchan * l3d;
#define XC_CREATE(call)                                                              \
  do {                                                                               \
    if ((call) != cudaSuccess) { cudaGetLastError(); gritas_b200_xcov_destroy(h); return GRITAS_B200_ERR_CUDA; } \
  } while (0)
  XC_CREATE(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  XC_CREATE(cudaMalloc(&h->d_lut, lut.size()));
  XC_CREATE(cudaMemcpy(h->d_lut, lut.data(), lut.size(), cudaMemcpyHostToDevice));
  XC_CREATE(cudaMalloc(&h->d_table, sizeof(int) * (size_t)kxmax * ktmax));
  XC_CREATE(cudaMemcpy(h->d_table, kxkt_table, sizeof(int) * (size_t)kxmax * ktmax, cudaMemcpyDefault));
  XC_CREATE(cudaMalloc(&h->bias, sizeof(double) * h->n_bias));
  XC_CREATE(cudaMalloc(&h->xcov, sizeof(double) * h->n_xcov));
  XC_CREATE(cudaMalloc(&h->nobs, sizeof(double) * h->n_nobs));
  XC_CREATE(cudaMalloc(&h->w_bias, sizeof(double) * h->n_bias));
  XC_CREATE(cudaMalloc(&h->w_xcov, sizeof(double) * h->n_xcov));
  XC_CREATE(cudaMalloc(&h->d_small, sizeof(int) * 4));
  XC_CREATE(cudaMemset(h->bias, 0, sizeof(double) * h->n_bias));
  XC_CREATE(cudaMemset(h->xcov, 0, sizeof(double) * h->n_xcov));
  XC_CREATE(cudaMemset(h->nobs, 0, sizeof(double) * h->n_nobs));
#undef XC_CREATE
  d.lut = h->d_lut;
  d.table = h->d_table;
  *hp = h;
  return 0;
}

void gritas_b200_xcov_destroy(gritas_b200_xcov_handle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->st) { cudaStreamSynchronize(h->st); cudaStreamDestroy(h->st); }
  for (void *p : {(void *)h->d_lut, (void *)h->d_table, (void *)h->bias, (void *)h->xcov, (void *)h->nobs, (void *)h->w_bias,
                  (void *)h->w_xcov, h->cols, (void *)h->accept, (void *)h->rank, (void *)h->X, (void *)h->L, (void *)h->d_small,
                  (void *)h->d_p1, (void *)h->d_p2, h->prs_scr, h->prs_tmp})
    if (p) cudaFree(p);
  cudaGetLastError();
  delete h;
}

const char *gritas_b200_xcov_last_error(gritas_b200_xcov_handle h) { return h ? h->err : "null handle"; }

int gritas_b200_xcov_reset(gritas_b200_xcov_handle h) {
  if (!h) return GRITAS_B200_ERR_ARG;
  XCU(cudaSetDevice(h->device));
  XCU(cudaMemsetAsync(h->bias, 0, sizeof(double) * h->n_bias, h->st));
  XCU(cudaMemsetAsync(h->xcov, 0, sizeof(double) * h->n_xcov, h->st));
  XCU(cudaMemsetAsync(h->nobs, 0, sizeof(double) * h->n_nobs, h->st));
  return 0;
}

int gritas_b200_xcov_accum(gritas_b200_xcov_handle h, int nobs, const double *lev, const int32_t *kx, const int32_t *kt,
                           const double *del, int32_t *ob_flag, int *nprof_out) {
  if (!h) return GRITAS_B200_ERR_ARG;
  if (h->prs) return xfail(h, GRITAS_B200_ERR_ARG, "this handle accumulates level covariances: use gritas_b200_xcov_prs_accum");
  XCU(cudaSetDevice(h->device));
  const XcDesc &d = h->d;
  if (nprof_out) *nprof_out = 0;
  if (nobs < 0) return xfail(h, GRITAS_B200_ERR_ARG, "nobs < 0");
  if (nobs % d.km != 0)                                                 // :751-757
    return xfail(h, GRITAS_B200_ERR_PROFILE, "XCov_Accum: not all profiles complete, aborting ... (nobs %d, profiles of %d)", nobs, d.km);
  if (nobs == 0) return 0;
  if (!lev || !kx || !kt || !del) return xfail(h, GRITAS_B200_ERR_ARG, "null column");
  const int nprof = nobs / d.km;
  const size_t n = (size_t)nobs;
  // columns onto the device (host arrays go through one scratch buffer; device arrays are used in place)
  const void *src[5] = {lev, kx, kt, del, ob_flag};
  const size_t bytes[5] = {8 * n, 4 * n, 4 * n, 8 * n * d.nr, 4 * n};
  size_t need = 0;
  for (int c = 0; c < 5; ++c)
    if (src[c] && !is_device_ptr(src[c])) need += (bytes[c] + 255) & ~(size_t)255;
  if (need) {
    int rc = xgrow(h, (char **)&h->cols, &h->cols_cap, need);
    if (rc) return rc;
  }
  const void *dev[5] = {};
  size_t off = 0;
  for (int c = 0; c < 5; ++c) {
    if (!src[c]) continue;
    if (is_device_ptr(src[c])) { dev[c] = src[c]; continue; }
    void *p = (char *)h->cols + off;
    XCU(cudaMemcpyAsync(p, src[c], bytes[c], cudaMemcpyHostToDevice, h->st));
    dev[c] = p;
    off += (bytes[c] + 255) & ~(size_t)255;
  }
  int rc;
  if ((rc = xgrow(h, &h->accept, &h->prof_cap, (size_t)nprof))) return rc;
  if ((rc = xgrow(h, &h->rank, &h->rank_cap, (size_t)nprof))) return rc;
  XCU(cudaMemsetAsync(h->d_small, 0, sizeof(int) * 4, h->st));
  const int wblocks = std::min(148 * 8, (nprof + 7) / 8);
  k_xc_scan<<<wblocks, 256, 0, h->st>>>(d, nprof, (const double *)dev[0], (const int *)dev[1], (const int *)dev[2], (int *)dev[4],
                                         h->accept, h->d_small + 1);
  k_xc_rank<<<1, 1024, 0, h->st>>>(h->accept, h->rank, nprof, h->d_small);
  int small[2] = {0, 0};
  XCU(cudaMemcpyAsync(small, h->d_small, sizeof(small), cudaMemcpyDeviceToHost, h->st));
  XCU(cudaStreamSynchronize(h->st));
  h->launches += 2;
  if (small[1]) return xfail(h, GRITAS_B200_ERR_PROFILE, "XCov-Accum: error in channel matching");       // :789-790
  const int nacc = small[0];
  if (ob_flag && !is_device_ptr(ob_flag)) XCU(cudaMemcpyAsync(ob_flag, dev[4], 4 * n, cudaMemcpyDeviceToHost, h->st));
  if (nacc > 0) {
    if ((rc = xgrow(h, &h->X, &h->x_cap, (size_t)nacc * d.nchan * d.nr))) return rc;
    if ((rc = xgrow(h, &h->L, &h->l_cap, (size_t)nacc * d.nchan))) return rc;
    k_xc_gather<<<wblocks, 256, 0, h->st>>>(d, nprof, nobs, (const double *)dev[0], (const int *)dev[1], (const int *)dev[2],
                                           (const double *)dev[3], h->accept, h->rank, nacc, h->X, h->L);
    const int nt = (d.nchan + XT - 1) / XT;
    dim3 grid(nt, nt, d.l3d * (d.tch ? 3 : 1));
    k_xc_outer<<<grid, 256, 0, h->st>>>(d, nacc, h->X, h->L, h->xcov);
    const int tot = d.nchan * d.l3d * (d.nr + 1);
    k_xc_bias<<<(tot + 255) / 256, 256, 0, h->st>>>(d, nacc, h->X, h->L, h->bias, h->nobs);
    h->launches += 3;
  }
  XCU(cudaGetLastError());
  XCU(cudaStreamSynchronize(h->st));
  if (nprof_out) *nprof_out = nacc;
  return 0;
}

// GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059).  The Norm, get_raw, reset and destroy entries are shared with the channel variant.
int gritas_b200_xcov_prs_create(gritas_b200_xcov_handle *hp, int nlevs, const double *p1, const double *p2, int l3d, int nr, int tch,
                                int device) {
  if (!hp) return GRITAS_B200_ERR_ARG;
  *hp = nullptr;
  if (nlevs < 1 || nlevs > 4096 || !p1 || !p2 || l3d < 1) return GRITAS_B200_ERR_ARG;
  std::vector<int32_t> achan(nlevs), table(1, 0);
  for (int k = 0; k < nlevs; ++k) achan[k] = k + 1;
  int rc = gritas_b200_xcov_create(hp, nlevs, nlevs, achan.data(), l3d, nr, tch, table.data(), 1, 1, device);
  if (rc) return rc;
  gritas_b200_xcov_handle h = *hp;
  h->prs = true;
  h->nlevs = nlevs;
  if (cudaMalloc(&h->d_p1, sizeof(double) * nlevs) != cudaSuccess || cudaMalloc(&h->d_p2, sizeof(double) * nlevs) != cudaSuccess ||
      cudaMemcpy(h->d_p1, p1, sizeof(double) * nlevs, cudaMemcpyDefault) != cudaSuccess ||
      cudaMemcpy(h->d_p2, p2, sizeof(double) * nlevs, cudaMemcpyDefault) != cudaSuccess) {
    cudaGetLastError();
    gritas_b200_xcov_destroy(h);
    *hp = nullptr;
    return GRITAS_B200_ERR_CUDA;
  }
  return 0;
}

int gritas_b200_xcov_prs_accum(gritas_b200_xcov_handle h, int nobs, const double *lev, const int32_t *ks, const double *del, int *nprof_out,
                               double *bad_lev) {
  if (!h) return GRITAS_B200_ERR_ARG;
  if (!h->prs) return xfail(h, GRITAS_B200_ERR_ARG, "this handle accumulates channel covariances: use gritas_b200_xcov_accum");
  XCU(cudaSetDevice(h->device));
  const XcDesc &d = h->d;
  if (nprof_out) *nprof_out = 0;
  if (nobs < 0) return xfail(h, GRITAS_B200_ERR_ARG, "nobs < 0");
  if (nobs == 0) return 0;
  if (!lev || !ks || !del) return xfail(h, GRITAS_B200_ERR_ARG, "null column");
  const size_t n = (size_t)nobs;
  const int nl = h->nlevs;
  // columns onto the device
  const void *src[3] = {lev, ks, del};
  const size_t bytes[3] = {8 * n, 4 * n, 8 * n * d.nr};
  size_t need = 0;
  for (int c = 0; c < 3; ++c)
    if (!is_device_ptr(src[c])) need += (bytes[c] + 255) & ~(size_t)255;
  int rc;
  if (need && (rc = xgrow(h, (char **)&h->cols, &h->cols_cap, need))) return rc;
  const void *dev[3] = {};
  size_t off = 0;
  for (int c = 0; c < 3; ++c) {
    if (is_device_ptr(src[c])) { dev[c] = src[c]; continue; }
    void *p = (char *)h->cols + off;
    XCU(cudaMemcpyAsync(p, src[c], bytes[c], cudaMemcpyHostToDevice, h->st));
    dev[c] = p;
    off += (bytes[c] + 255) & ~(size_t)255;
  }
  // scratch: nine int-sized columns of n
  const size_t col = (4 * n + 255) & ~(size_t)255;
  if ((rc = xgrow(h, (char **)&h->prs_scr, &h->prs_cap, 10 * col + 256))) return rc;
  char *scr = (char *)h->prs_scr;
  int *kidx = (int *)(scr + 0 * col);
  unsigned *start = (unsigned *)(scr + 1 * col), *srank = (unsigned *)(scr + 2 * col), *key = (unsigned *)(scr + 3 * col);
  int *idx = (int *)(scr + 4 * col);
  unsigned *skey = (unsigned *)(scr + 5 * col);
  int *sidx = (int *)(scr + 6 * col);
  unsigned *gstart = (unsigned *)(scr + 7 * col), *grank = (unsigned *)(scr + 8 * col);
  int *lm = (int *)(scr + 9 * col);
  const int nb = (int)std::min<size_t>(148 * 8, (n + 255) / 256);
  const int big = 0x7fffffff;
  XCU(cudaMemcpyAsync(h->d_small, &big, sizeof(int), cudaMemcpyHostToDevice, h->st));
  k_xp_levels<<<nb, 256, 0, h->st>>>(nobs, nl, (const double *)dev[0], h->d_p1, h->d_p2, kidx, h->d_small);
  k_xp_runs<<<nb, 256, 0, h->st>>>(nobs, (const int *)dev[1], start, key, idx);
  k_xc_rank<<<1, 1024, 0, h->st>>>(start, srank, nobs, h->d_small + 1);
  size_t tmpb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmpb, key, skey, idx, sidx, nobs, 0, 32, h->st);
  if ((rc = xgrow(h, (char **)&h->prs_tmp, &h->prs_tmp_cap, tmpb + 256))) return rc;
  size_t tb = h->prs_tmp_cap;
  XCU(cub::DeviceRadixSort::SortPairs(h->prs_tmp, tb, key, skey, idx, sidx, nobs, 0, 32, h->st));   // stable: members keep their input order
  k_xp_gstart<<<nb, 256, 0, h->st>>>(nobs, skey, gstart);
  k_xc_rank<<<1, 1024, 0, h->st>>>(gstart, grank, nobs, h->d_small + 2);
  int small[3] = {0, 0, 0};
  XCU(cudaMemcpyAsync(small, h->d_small, sizeof(small), cudaMemcpyDeviceToHost, h->st));
  XCU(cudaStreamSynchronize(h->st));
  h->launches += 6;
  if (small[0] != big) {                                               // 'Accum: could not find level(1)' (:1990)
    double lv = 0.0;
    XCU(cudaMemcpy(&lv, (const double *)dev[0] + small[0], sizeof(double), cudaMemcpyDefault));
    if (bad_lev) *bad_lev = lv;
    return xfail(h, GRITAS_B200_ERR_LEVEL, "Accum: could not find level(1) (obs level %.17g hPa)", lv);
  }
  const int nruns = small[1], ngroups = small[2];
  // run keys, group begins, level buckets, run -> group
  const size_t g_off = (size_t)ngroups * (nl + 1);
  if ((rc = xgrow(h, &h->rank, &h->rank_cap, (size_t)nruns * 2 + (size_t)ngroups + 2 + g_off))) return rc;
  unsigned *run_key = h->rank, *gbeg = run_key + nruns, *goff = gbeg + ngroups + 1;
  int *run_group = (int *)(goff + g_off);
  k_xp_run_ks<<<nb, 256, 0, h->st>>>(nobs, (const int *)dev[1], start, srank, run_key);
  k_xp_groups<<<nb, 256, 0, h->st>>>(nobs, gstart, grank, ngroups, gbeg);
  k_xp_bucket<<<(ngroups + 255) / 256, 256, 0, h->st>>>(nl, sidx, kidx, ngroups, gbeg, goff, lm);
  k_xp_run_group<<<(nruns + 255) / 256, 256, 0, h->st>>>(nruns, ngroups, run_key, skey, gbeg, run_group);
  const int cells = nl * nl * (d.tch ? d.nr : 1);
  k_xp_accum<<<(cells + 127) / 128, 128, 0, h->st>>>(nl, d.l3d, d.nr, d.tch, nobs, nruns, run_group, goff, lm, (const double *)dev[2], h->bias,
                                                    h->xcov, h->nobs);
  h->launches += 5;
  XCU(cudaGetLastError());
  XCU(cudaStreamSynchronize(h->st));
  if (nprof_out) *nprof_out = nruns;
  return 0;
}

static int xc_download(gritas_b200_xcov_handle h, const double *db, const double *dx, double *bias_lv, double *xcov_lv, double *nobs_lv) {
  if (bias_lv) XCU(cudaMemcpyAsync(bias_lv, db, sizeof(double) * h->n_bias, cudaMemcpyDefault, h->st));
  if (xcov_lv) XCU(cudaMemcpyAsync(xcov_lv, dx, sizeof(double) * h->n_xcov, cudaMemcpyDefault, h->st));
  if (nobs_lv) XCU(cudaMemcpyAsync(nobs_lv, h->nobs, sizeof(double) * h->n_nobs, cudaMemcpyDefault, h->st));
  XCU(cudaStreamSynchronize(h->st));
  return 0;
}

int gritas_b200_xcov_get_raw(gritas_b200_xcov_handle h, double *bias_lv, double *xcov_lv, double *nobs_lv) {
  if (!h) return GRITAS_B200_ERR_ARG;
  XCU(cudaSetDevice(h->device));
  return xc_download(h, h->bias, h->xcov, bias_lv, xcov_lv, nobs_lv);
}

int gritas_b200_xcov_norm(gritas_b200_xcov_handle h, int nrmlz, int self, int ntresh, double *bias_lv, double *xcov_lv,
                          double *nobs_lv) {
  if (!h) return GRITAS_B200_ERR_ARG;
  XCU(cudaSetDevice(h->device));
  const XcDesc &d = h->d;
  const int ntr = ntresh < 0 ? 0 : ntresh;                              // absent: no effect (:931-934)
  XCU(cudaMemcpyAsync(h->w_bias, h->bias, sizeof(double) * h->n_bias, cudaMemcpyDeviceToDevice, h->st));
  XCU(cudaMemcpyAsync(h->w_xcov, h->xcov, sizeof(double) * h->n_xcov, cudaMemcpyDeviceToDevice, h->st));
  k_xc_norm_bias<<<(int)std::min<size_t>(148 * 8, (h->n_bias + 255) / 256), 256, 0, h->st>>>(d, nrmlz ? 1 : 0, ntr, h->nobs, h->w_bias);
  const size_t plane = (size_t)d.nchan * d.nchan * d.l3d;
  k_xc_norm_cov<<<(int)std::min<size_t>(148 * 8, (plane + 255) / 256), 256, 0, h->st>>>(d, nrmlz ? 1 : 0, ntr, self ? 1 : 0, h->nobs,
                                                                                         h->w_bias, h->w_xcov);
  h->launches += 2;
  XCU(cudaGetLastError());
  return xc_download(h, h->w_bias, h->w_xcov, bias_lv, xcov_lv, nobs_lv);
}

uint64_t gritas_b200_xcov_launch_count(gritas_b200_xcov_handle h) { return h ? h->launches : 0; }

}  // extern "C"
