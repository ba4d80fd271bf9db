/*
 * gritas_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the GrITAS gridding hot path.
 *
 * This file is the checker ("oracle") for gritas_b200.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (gritas_b200/, include/) never links against or falls back to anything in oracle/.
 *
 * PARITY: the reference (GEOS-ESM/GrITAS) ships no tests, golden vectors or sample data for this path and cannot be
 * built here (no Fortran compiler; GMAO_Shared / ESMA_cmake are un-vendored externals).  What this restatement is
 * pinned against instead:
 *   - PINNED against the reference's own source text: oracle/f2c_subset.py translates to C, STATEMENT BY STATEMENT and from
 *     the files where they lie: GrITAS_Pint_, GrITAS_Accum_, GrITAS_Norm_, GrITAS_MinMax_, GrITAS_3CH_Norm_,
 *     GrITAS_XCov_Accum_, GrITAS_XCov_Norm_, GrITAS_XCov_Prs_Accum_, scores2_, ave_nomiss, the mesh statements and the region rows of
 *     init_ (m_gritas_grids.f), GrITAS_KxKt (gritas_kxkt.f), the loop of maskout_ (m_gritas_masks.F90), the pre-sort call
 *     sequence, the residual selection chain, the QC / time / longitude filter loop and the Jo scaling (gritas.f:500-508,
 *     562-704, 712-735, 738).  gcc compiles the result (oracle/_ref, git-ignored) and tests/test_oracle_vs_ref.py (119
 *     cases) demands bit-for-bit agreement with the routines below -- flags, sums, statistics, tables, permutations,
 *     error statements -- on random inputs, the committed golden fixtures and the edge cases of SURVEY.md section 4;
 *   - ASSUMED, because their sources are not in the reference tree: a stable merge sort for m_MergeSorts' IndexSort
 *     (GMAO_mpeu), X_PASSIVE = 7 / X_PRE_BAD = 1 / ktSurfAll of m_odsmeta (GMAO_ods; parameters everywhere), the ODS file
 *     format.
 *   - one DEFECT OF THE REFERENCE found this way and not reproduced: GrITAS_XCov_Prs_Accum_ books the first observation of
 *     a sounding to the sounding before it (:1961-1975), allocates the last sounding of a call one element short and
 *     gathers past the end of its work arrays (undefined behaviour); the oracle and the CUDA path compute what the loop
 *     intends.
 * The restatement follows the reference source literally, one routine per reference routine.  All citations are
 * relative to /root/reference/.
 *
 * Conventions kept from the reference build:
 *   - every `real` is 64-bit (FREAL8: src/Components/gritas/CMakeLists.txt:3-4), `integer` is
 *     32-bit;
 *   - arrays are column-major and 1-based (macros below);
 *   - Fortran NINT rounds half away from zero (lround), not to even;
 *   - compile with -ffp-contract=off so `rtms + del*del` is a multiply then an add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_AMISS 1.0e15  /* m_gritas_grids.f:72 */
#define ORC_LONMIN (-180.0) /* m_gritas_grids.f:33 */
#define ORC_LATMIN (-90.0)  /* m_gritas_grids.f:42 */

/* Fortran NINT for a 64-bit real into a default integer.  In range this is lround(); out of
 * range (and NaN) is undefined in Fortran -- we pin it to saturation so the result is the same on
 * every platform (NaN -> most negative, like the x86 "integer indefinite"). */
static long orc_nint(double x) {
  if (!(x == x)) return (long)INT32_MIN;
  if (x >= 2147483647.0) return (long)INT32_MAX;
  if (x <= -2147483648.0) return (long)INT32_MIN;
  return lround(x);
}

/* m_gritas_grids.f:128-129 (init_): mesh sizes from the grid dimensions. */
void orc_grid_init(int im, int jm, double *dlon, double *dlat) {
  *dlon = 360.0 / (double)im;
  *dlat = 180.0 / (double)(jm - 1);
}

/* m_gritas_grids.f:217-296 (GrITAS_Pint_).  Returns 0, or the gritas_perror case:
 * 1 = nlevs<=0, 2 = not monotonically decreasing, 3 = last level is 0 hPa. */
int orc_pint(int nlevs, const double *plevs, double *p1, double *p2) {
  int k;
  if (nlevs <= 0) return 1;                       /* :244-245 */
  if (nlevs == 1) {                               /* :249-256 */
    p1[0] = plevs[0];
    p2[0] = plevs[0] - 0.1;
    return 0;
  }
  for (k = 0; k < nlevs - 1; ++k)                 /* :261-264 */
    if (plevs[k] <= plevs[k + 1]) return 2;
  if (plevs[nlevs - 1] == 0.0) return 3;          /* :266-267 */
  p1[0] = 2000.0;                                 /* :270 */
  p2[0] = exp(0.5 * (log(plevs[1]) + log(plevs[0])));              /* :271 */
  for (k = 1; k < nlevs - 1; ++k) {               /* :273-276 */
    p1[k] = exp(0.5 * (log(plevs[k - 1]) + log(plevs[k])));
    p2[k] = exp(0.5 * (log(plevs[k + 1]) + log(plevs[k])));
  }
  k = nlevs - 1;                                  /* :278-280 */
  p1[k] = exp(0.5 * (log(plevs[k - 1]) + log(plevs[k])));
  p2[k] = 0.0;
  return 0;
}

/* gritas_kxkt.f:133-164 (is_surface): kt<0 or kt in ktSurfAll (m_odsmeta, not in tree: the
 * list is passed in). */
static int orc_is_surface(int kt, const int *kt_surf, int n_surf) {
  int i;
  if (kt < 0) return 1;
  for (i = 0; i < n_surf; ++i)
    if (kt_surf[i] == kt) return 1;
  return 0;
}

/* gritas_kxkt.f:10-119 (GrITAS_KxKt).  kx_list is (kxmax, listsz) column-major, terminated by a
 * negative entry per column.  kxkt_table is (kxmax, ktmax) column-major.
 * Returns 0 or: 1 kt too large, 2 kt<1, 3 l2d too large, 4 l3d too large, 5 kx too large.
 * A kx of 0 would index out of bounds in the reference; it is skipped here. */
int orc_kxkt(int ktmax, int kxmax, int l2d_max, int l3d_max, int listsz, const int *kt_list,
             const int *kx_list, const int *kt_surf, int n_surf, int *kxkt_table, int *list_2d,
             int *list_3d, int *l2d_out, int *l3d_out) {
  int i, j, kx, kt, l, l2d = 0, l3d = 0;
  memset(kxkt_table, 0, sizeof(int) * (size_t)kxmax * (size_t)ktmax);   /* :73-77 */
  for (j = 1; j <= listsz; ++j) {                                        /* :81 */
    kt = abs(kt_list[j - 1]);
    if (kt > ktmax) return 1;
    if (kt < 1) return 2;
    if (orc_is_surface(kt_list[j - 1], kt_surf, n_surf)) {               /* :87-92 */
      l2d++;
      if (l2d > l2d_max) return 3;
      l = -l2d;
      list_2d[l2d - 1] = j;
    } else {                                                             /* :93-98 */
      l3d++;
      if (l3d > l3d_max) return 4;
      l = l3d;
      list_3d[l3d - 1] = j;
    }
    for (i = 1; i <= kxmax; ++i) {                                       /* :100-106 */
      kx = kx_list[(size_t)(i - 1) + (size_t)kxmax * (size_t)(j - 1)];
      if (kx < 0) break;
      if (kx > kxmax) return 5;
      if (kx == 0) continue;
      kxkt_table[(size_t)(kx - 1) + (size_t)kxmax * (size_t)(kt - 1)] = l; /* last writer wins */
    }
  }
  *l2d_out = l2d;
  *l3d_out = l3d;
  return 0;
}

/* Stable index merge sort on a double key (the role of m_MergeSorts::IndexSort, GMAO_mpeu --
 * not in tree; gritas.f:500-508 only relies on "stable, ascending").  `is` holds 0-based
 * indices into key. */
static void orc_index_sort(int n, int *is, const double *key, int *tmp) {
  int width, lo;
  for (width = 1; width < n; width *= 2) {
    for (lo = 0; lo < n; lo += 2 * width) {
      int mid = lo + width < n ? lo + width : n;
      int hi = lo + 2 * width < n ? lo + 2 * width : n;
      int a = lo, b = mid, o = lo;
      while (a < mid && b < hi) {
        if (key[is[b]] < key[is[a]]) tmp[o++] = is[b++]; /* strict: ties keep left first */
        else tmp[o++] = is[a++];
      }
      while (a < mid) tmp[o++] = is[a++];
      while (b < hi) tmp[o++] = is[b++];
    }
    memcpy(is, tmp, sizeof(int) * (size_t)n);
  }
}

/* gritas.f:499-508: IndexSet then 8 successive stable ascending sorts
 * (qcexcl, lat, lon, lev, time, kt, ks, kx).  Output `is` is 0-based. */
int orc_presort(int nobs, const int *qcexcl, const double *lat, const double *lon,
                const double *lev, const int *time, const int *kt, const int *ks, const int *kx,
                int *is) {
  int i, p;
  int *tmp = (int *)malloc(sizeof(int) * (size_t)(nobs > 0 ? nobs : 1));
  double *key = (double *)malloc(sizeof(double) * (size_t)(nobs > 0 ? nobs : 1));
  if (!tmp || !key) { free(tmp); free(key); return 1; }
  for (i = 0; i < nobs; ++i) is[i] = i;                                  /* IndexSet */
  for (p = 0; p < 8; ++p) {
    const int *ik = NULL;
    const double *dk = NULL;
    switch (p) {
      case 0: ik = qcexcl; break;
      case 1: dk = lat; break;
      case 2: dk = lon; break;
      case 3: dk = lev; break;
      case 4: ik = time; break;
      case 5: ik = kt; break;
      case 6: ik = ks; break;
      default: ik = kx; break;
    }
    if (ik) { for (i = 0; i < nobs; ++i) key[i] = (double)ik[i]; dk = key; }
    orc_index_sort(nobs, is, dk, tmp);
  }
  free(tmp);
  free(key);
  return 0;
}

/* gritas.f:510-523: column gathers through the sort index. */
void orc_gather_f64(int nobs, const int *is, const double *src, double *dst) {
  int i;
  for (i = 0; i < nobs; ++i) dst[i] = src[is[i]];
}
void orc_gather_i32(int nobs, const int *is, const int *src, int *dst) {
  int i;
  for (i = 0; i < nobs; ++i) dst[i] = src[is[i]];
}

/* gritas.f:562-587: residual selection for iopt 0/1 (omf) and 2/3 (oma), with the optional
 * scale_res division.  |iopt| is used.  del is (nobs, nr) column-major; only column 1 is set
 * here (column 2 is set by the caller in the omf+oma one-pass mode, which is two reference
 * runs fused).  Returns 1 for an iopt this restatement does not cover. */
int orc_select_del(int nobs, int iopt, int scale_res, const double *omf, const double *oma,
                   const double *obs, const double *xvec, double *del) {
  int i, a = abs(iopt);
  const double *src;
  if (a == 0 || a == 1) src = omf;
  else if (a == 2 || a == 3) src = oma;
  else return 1;
  for (i = 0; i < nobs; ++i) {
    double d = src[i];
    if (scale_res == 1) d = d / sqrt(xvec[i]);
    if (scale_res == 2) d = d / obs[i];
    if (scale_res == 3) d = d / (obs[i] - omf[i]);
    del[i] = d;
  }
  return 0;
}

/* gritas.f:562-704, every residual option that works on one ODS structure (no -meanres second file):
 * fills del (nobs, nr) column-major and returns nr (1..3), or 0 for an option this restatement does not
 * cover (26/27, invalid).  Includes the division of gritas.f:738 (iopt 17/19: del(:,1)/del(:,2)), which the
 * reference performs right after the QC filter.  `amiss` is m_gritas_grids' AMISS. */
int orc_select_del_all(int nobs, int iopt, int tch, int self, int scale_res, const double *omf, const double *oma,
                       const double *obs, const double *xvec, const double *xm, double *del) {
  int i, a = abs(iopt), nr = 1;
  double *d1 = del, *d2 = del + (size_t)nobs, *d3 = del + 2 * (size_t)nobs;
  if (a == 0 || a == 1 || a == 2 || a == 3) {                         /* :562-587 */
    if (orc_select_del(nobs, iopt, scale_res, omf, oma, obs, xvec, del)) return 0;
    return 1;
  }
  for (i = 0; i < nobs; ++i) {
    switch (a) {
      case 4: case 5: d1[i] = obs[i]; break;                          /* :588-589 */
      case 6: case 7: d1[i] = xvec[i]; break;                         /* :590-591 */
      case 8: case 9: d1[i] = xm[i]; break;                           /* :592-593 */
      case 10: case 11:                                               /* :594-611 */
        if (tch) { d1[i] = oma[i] - omf[i]; d2[i] = -omf[i]; d3[i] = -oma[i]; nr = 3; }
        else { d1[i] = omf[i] - oma[i]; d2[i] = omf[i]; nr = 2; }
        break;
      case 12: case 13:                                               /* :612-624 */
        if (tch) { d1[i] = -oma[i]; d2[i] = omf[i] - oma[i]; d3[i] = omf[i]; nr = 3; }
        else { d1[i] = omf[i] - oma[i]; d2[i] = oma[i]; nr = 2; }
        break;
      case 14: case 15:                                               /* :625-640 */
        if (tch && self) { d1[i] = obs[i]; d2[i] = obs[i] - omf[i]; d3[i] = obs[i] - oma[i]; nr = 3; }
        else if (tch) { d1[i] = omf[i]; d2[i] = oma[i]; d3[i] = oma[i] - omf[i]; nr = 3; }
        else { d1[i] = oma[i]; d2[i] = omf[i]; nr = 2; }
        break;
      case 16: case 17: d1[i] = omf[i]; d2[i] = xvec[i]; nr = 2; break; /* :641-644 */
      case 18: case 19: d1[i] = oma[i]; d2[i] = xvec[i]; nr = 2; break; /* :645-648 */
      case 20: case 21: d1[i] = obs[i] + xm[i]; break;                /* :649-650 */
      case 22: case 23: d1[i] = xvec[i]; break;                       /* :651-652 */
      case 24: case 25: d1[i] = obs[i] - omf[i]; break;               /* :663 (not -meanres) */
      case 28: case 29: d1[i] = obs[i] - omf[i] + xm[i]; break;       /* :678-679 */
      case 30: case 31: d1[i] = obs[i] - oma[i]; break;               /* :690 (not -meanres) */
      case 32: case 33:                                               /* :692-701 */
        d1[i] = omf[i] - oma[i];
        d2[i] = oma[i];
        if (xvec[i] > 0.0) { d1[i] = d1[i] / xvec[i]; d2[i] = d2[i] / xvec[i]; }
        else { d1[i] = ORC_AMISS; d2[i] = ORC_AMISS; }
        nr = 2;
        break;
      default: return 0;
    }
    if (a == 17 || a == 19) d1[i] = d1[i] / d2[i];                    /* :738, inside the odd-iopt filter block */
  }
  return nr;
}

/* gritas.f:710-739: QC / time / longitude filter producing ob_flag (0 keep, 1 drop).  The
 * caller only runs it when iopt is odd (always the case from the CLI). */
void orc_filter(int nobs, const int *qcexcl, const int *time, const double *lon, int ioptn,
                int t1, int t2, double lona, double lonb, int x_passive, int x_pre_bad,
                int *ob_flag) {
  int i;
  for (i = 0; i < nobs; ++i) {
    int iampassive = qcexcl[i] == x_passive || qcexcl[i] == x_pre_bad;
    int iwantpassive = ioptn > 0 && iampassive;
    int timeselect, lonselect;
    if (t2 < t1) timeselect = 1;
    else if (t1 <= time[i] && time[i] <= t2) timeselect = 1;
    else timeselect = 0;
    if (lona < -180.0 && lonb > 180.0) lonselect = 1;
    else if (lona <= lon[i] && lon[i] <= lonb) lonselect = 1;
    else lonselect = 0;
    if (timeselect && lonselect && (qcexcl[i] == 0 || iwantpassive)) ob_flag[i] = 0;
    else ob_flag[i] = 1;
  }
}

/* Column-major 1-based offsets (m_gritas_grids.f:346-355). */
#define IX2(i, j, l, ir) \
  ((size_t)((i)-1) + (size_t)im * ((size_t)((j)-1) + (size_t)jm * ((size_t)((l)-1) + (size_t)l2d * (size_t)((ir)-1))))
#define IX3(i, j, k, l, ir)                                                               \
  ((size_t)((i)-1) +                                                                      \
   (size_t)im * ((size_t)((j)-1) +                                                        \
                 (size_t)jm * ((size_t)((k)-1) +                                          \
                               (size_t)km * ((size_t)((l)-1) + (size_t)l3d * (size_t)((ir)-1)))))

/* m_gritas_grids.f:307-478 (GrITAS_Accum_).  del is (nobs, nr) column-major; kxkt_table is
 * (kxmax, ktmax) column-major.  xcov_* are addressed as (im,jm,l2d) / (im,jm,km,l3d), i.e. slot 1
 * of the caller's 4-D/5-D array (gritas.f:780-781).
 *
 * Returns 0; 7 on "could not find level" (the reference prints the level and exits with 7,
 * m_gritas_grids.f:458-459 -> gritas_etc.f:54-60): *bad_lev holds the level and *bad_n the
 * 0-based observation, and accumulation stops there exactly as the reference's would;
 * 8 for nr>3 (:395-396).
 * A (kx,kt) outside the table is undefined behaviour in the reference (:409, no bounds check);
 * here it is treated as "not in table". */
int orc_accum(int im, int jm, int km, int nlevs, int nobs, int nr, const double *lat,
              const double *lon, const double *lev, const int *kx, const int *kt,
              const double *del, const int *kxkt_table, int kxmax, int ktmax, const double *p1,
              const double *p2, int *ob_flag, int l2d, double *bias_2d, double *rtms_2d,
              double *xcov_2d, double *nobs_2d, int l3d, double *bias_3d, double *rtms_3d,
              double *xcov_3d, double *nobs_3d, double *bad_lev, int *bad_n) {
  int n, i, j, k, kk, nx, l, ir, surface_var;
  double dlon, dlat;
  long ii, jj;
  orc_grid_init(im, jm, &dlon, &dlat);
  if (nr > 3) return 8;                                                  /* :395-396 */
  nx = (nr == 3) ? 0 : (nr > 1 ? nr : 1);                                /* :397-401 */

  for (n = 0; n < nobs; ++n) {                                           /* :405 */
    if (ob_flag[n] != 0) continue;                                       /* :406-407 */
    if (kx[n] < 1 || kx[n] > kxmax || kt[n] < 1 || kt[n] > ktmax) l = 0;
    else l = kxkt_table[(size_t)(kx[n] - 1) + (size_t)kxmax * (size_t)(kt[n] - 1)]; /* :409 */
    surface_var = l < 0;                                                 /* :410 */
    l = abs(l);                                                          /* :411 */
    if (l <= 0) {                                                        /* :413-416 */
      ob_flag[n] = 1;
      continue;
    }
    ii = 1 + orc_nint((lon[n] - ORC_LONMIN) / dlon);                     /* :420 */
    if (ii == (long)im + 1) ii = 1;                                      /* :421 */
    if (ii == 0) ii = im;                                                /* :422 */
    jj = 1 + orc_nint((lat[n] - ORC_LATMIN) / dlat);                     /* :423 */
    if (ii < 1) ii = 1;                                                  /* :425 */
    if (ii > im) ii = im;
    if (jj < 1) jj = 1;                                                  /* :426 */
    if (jj > jm) jj = jm;
    i = (int)ii;
    j = (int)jj;

    if (surface_var) {                                                   /* :432-441 */
      for (ir = 1; ir <= nr; ++ir) {
        double d = del[(size_t)n + (size_t)nobs * (size_t)(ir - 1)];
        bias_2d[IX2(i, j, l, ir)] = bias_2d[IX2(i, j, l, ir)] + d;
        rtms_2d[IX2(i, j, l, ir)] = rtms_2d[IX2(i, j, l, ir)] + d * d;
      }
      if (nx > 0)
        xcov_2d[IX2(i, j, l, 1)] =
            xcov_2d[IX2(i, j, l, 1)] + del[n] * del[(size_t)n + (size_t)nobs * (size_t)(nx - 1)];
      nobs_2d[IX2(i, j, l, 1)] = nobs_2d[IX2(i, j, l, 1)] + 1.0;
    } else {
      k = 0;
      for (kk = 1; kk <= nlevs; ++kk) {                                  /* :450-456 */
        if (lev[n] <= p1[kk - 1] && lev[n] > p2[kk - 1]) {
          k = kk;
          break;
        }
      }
      if (k == 0) {                                                      /* :458-459 */
        if (bad_lev) *bad_lev = lev[n];
        if (bad_n) *bad_n = n;
        return 7;
      }
      for (ir = 1; ir <= nr; ++ir) {                                     /* :465-468 */
        double d = del[(size_t)n + (size_t)nobs * (size_t)(ir - 1)];
        bias_3d[IX3(i, j, k, l, ir)] = bias_3d[IX3(i, j, k, l, ir)] + d;
        rtms_3d[IX3(i, j, k, l, ir)] = rtms_3d[IX3(i, j, k, l, ir)] + d * d;
      }
      if (nx > 0)                                                        /* :469 */
        xcov_3d[IX3(i, j, k, l, 1)] =
            xcov_3d[IX3(i, j, k, l, 1)] + del[n] * del[(size_t)n + (size_t)nobs * (size_t)(nx - 1)];
      nobs_3d[IX3(i, j, k, l, 1)] = nobs_3d[IX3(i, j, k, l, 1)] + 1.0;   /* :470 */
    }
  }
  return 0;
}

/* Box assignment only (the K1 half of Accum, m_gritas_grids.f:405-461), for index parity tests:
 * key[n] = -1 rejected on entry, -2 not in table, -3 level miss, else the 0-based column-major
 * offset into nobs_3d (upper air) or -(4 + offset into nobs_2d) (surface). */
void orc_box_keys(int im, int jm, int km, int nlevs, int nobs, const double *lat,
                  const double *lon, const double *lev, const int *kx, const int *kt,
                  const int *kxkt_table, int kxmax, int ktmax, const double *p1, const double *p2,
                  const int *ob_flag, int l2d, int l3d, int64_t *key) {
  int n, kk, k, l, surface_var;
  long ii, jj;
  double dlon, dlat;
  (void)l2d; (void)l3d;
  orc_grid_init(im, jm, &dlon, &dlat);
  for (n = 0; n < nobs; ++n) {
    if (ob_flag[n] != 0) { key[n] = -1; continue; }
    if (kx[n] < 1 || kx[n] > kxmax || kt[n] < 1 || kt[n] > ktmax) l = 0;
    else l = kxkt_table[(size_t)(kx[n] - 1) + (size_t)kxmax * (size_t)(kt[n] - 1)];
    surface_var = l < 0;
    l = abs(l);
    if (l <= 0) { key[n] = -2; continue; }
    ii = 1 + orc_nint((lon[n] - ORC_LONMIN) / dlon);
    if (ii == (long)im + 1) ii = 1;
    if (ii == 0) ii = im;
    jj = 1 + orc_nint((lat[n] - ORC_LATMIN) / dlat);
    if (ii < 1) ii = 1;
    if (ii > im) ii = im;
    if (jj < 1) jj = 1;
    if (jj > jm) jj = jm;
    if (surface_var) {
      key[n] = -(4 + (int64_t)((ii - 1) + (int64_t)im * ((jj - 1) + (int64_t)jm * (l - 1))));
    } else {
      k = 0;
      for (kk = 1; kk <= nlevs; ++kk)
        if (lev[n] <= p1[kk - 1] && lev[n] > p2[kk - 1]) { k = kk; break; }
      if (k == 0) { key[n] = -3; continue; }
      key[n] = (int64_t)(ii - 1) +
               (int64_t)im * ((jj - 1) + (int64_t)jm * ((k - 1) + (int64_t)km * (l - 1)));
    }
  }
}

/* m_gritas_grids.f:490-656 (GrITAS_Norm_).  xcov_* here are the full (…, nr) arrays.
 * `m` is an INTEGER in the reference (:548): m = n when nrmlz, else stays 1, so that
 * m*tmp/(m-1) divides by zero under -nonorm exactly as the reference does.
 * `tmp`/`xtmp` are locals that persist across boxes in the reference; in the 3-D n==1 branch
 * (:642-644) with ntresh>1 they are read without having been set for this box, so the stale
 * values of the last box that passed the threshold are used (uninitialised before the first
 * such box: pinned to 0 here).  Returns 8 for nr>2 (:555). */
int orc_norm(int im, int jm, int km, int nlevs, int nr, int nrmlz, int l2d, double *bias_2d,
             double *rtms_2d, double *stdv_2d, double *xcov_2d, double *nobs_2d, int l3d,
             double *bias_3d, double *rtms_3d, double *stdv_3d, double *xcov_3d, double *nobs_3d,
             int ntresh) {
  int i, j, k, l, m, nx, ir;
  double tmp[2] = {0.0, 0.0}, xtmp = 0.0, n;
  if (nr > 2) return 8;
  nx = nr > 1 ? nr : 1;                                                  /* :556 */

  m = 1;                                                                 /* :566 */
  for (l = 1; l <= l2d; ++l)
    for (j = 1; j <= jm; ++j)
      for (i = 1; i <= im; ++i) {
        n = nobs_2d[IX2(i, j, l, 1)];                                    /* :571 */
        if (nrmlz) m = (int)n;                                           /* :572 */
        if (n > 0.0 && n >= (double)ntresh) {                            /* :573-582 */
          for (ir = 1; ir <= nr; ++ir) {
            bias_2d[IX2(i, j, l, ir)] = bias_2d[IX2(i, j, l, ir)] / (double)m;
            tmp[ir - 1] = rtms_2d[IX2(i, j, l, ir)] / (double)m;
            rtms_2d[IX2(i, j, l, ir)] = sqrt(tmp[ir - 1]);
            tmp[ir - 1] = fabs(tmp[ir - 1] - bias_2d[IX2(i, j, l, ir)] * bias_2d[IX2(i, j, l, ir)]);
          }
          xtmp = xcov_2d[IX2(i, j, l, 1)] / (double)m;
          xcov_2d[IX2(i, j, l, 1)] = xtmp;
          xtmp = xtmp - bias_2d[IX2(i, j, l, 1)] * bias_2d[IX2(i, j, l, nx)];
        } else {                                                         /* :583-589 */
          for (ir = 1; ir <= nr; ++ir) {
            bias_2d[IX2(i, j, l, ir)] = ORC_AMISS;
            rtms_2d[IX2(i, j, l, ir)] = ORC_AMISS;
          }
          xcov_2d[IX2(i, j, l, 1)] = ORC_AMISS;
        }
        if (n > 1.0 && n >= (double)ntresh) {                            /* :591-596 */
          for (ir = 1; ir <= nr; ++ir)
            stdv_2d[IX2(i, j, l, ir)] = sqrt((double)m * tmp[ir - 1] / (double)(m - 1));
          xcov_2d[IX2(i, j, l, nx)] = (double)m * xtmp / (double)(m - 1);
        } else {                                                         /* :597-601 */
          for (ir = 1; ir <= nr; ++ir) stdv_2d[IX2(i, j, l, ir)] = ORC_AMISS;
          xcov_2d[IX2(i, j, l, nx)] = ORC_AMISS;
        }
      }

  if (nlevs < 2) return 0;                                               /* :607 */

  m = 1;                                                                 /* :612 */
  for (l = 1; l <= l3d; ++l)
    for (k = 1; k <= nlevs; ++k)
      for (j = 1; j <= jm; ++j)
        for (i = 1; i <= im; ++i) {
          n = nobs_3d[IX3(i, j, k, l, 1)];                               /* :618 */
          if (nrmlz) m = (int)n;                                         /* :619 */
          if (n > 0.0 && n >= (double)ntresh) {                          /* :620-629 */
            for (ir = 1; ir <= nr; ++ir) {
              bias_3d[IX3(i, j, k, l, ir)] = bias_3d[IX3(i, j, k, l, ir)] / (double)m;
              tmp[ir - 1] = rtms_3d[IX3(i, j, k, l, ir)] / (double)m;
              rtms_3d[IX3(i, j, k, l, ir)] = sqrt(tmp[ir - 1]);
              tmp[ir - 1] =
                  fabs(tmp[ir - 1] - bias_3d[IX3(i, j, k, l, ir)] * bias_3d[IX3(i, j, k, l, ir)]);
            }
            xtmp = xcov_3d[IX3(i, j, k, l, 1)] / (double)m;
            xcov_3d[IX3(i, j, k, l, 1)] = xtmp;
            xtmp = xtmp - bias_3d[IX3(i, j, k, l, 1)] * bias_3d[IX3(i, j, k, l, nx)];
          } else {                                                       /* :630-636 */
            for (ir = 1; ir <= nr; ++ir) {
              bias_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
              rtms_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
            }
            xcov_3d[IX3(i, j, k, l, 1)] = ORC_AMISS;
          }
          if (n > 1.0 && n >= (double)ntresh) {                          /* :638-640 */
            for (ir = 1; ir <= nr; ++ir)
              stdv_3d[IX3(i, j, k, l, ir)] = sqrt((double)m * tmp[ir - 1] / (double)(m - 1));
            xcov_3d[IX3(i, j, k, l, nx)] = (double)m * xtmp / (double)(m - 1);
          } else if (n == 1.0) {                                         /* :642-644 */
            for (ir = 1; ir <= nr; ++ir) stdv_3d[IX3(i, j, k, l, ir)] = sqrt(tmp[ir - 1]);
            xcov_3d[IX3(i, j, k, l, nx)] = xtmp;
          } else {                                                       /* :646-647 */
            for (ir = 1; ir <= nr; ++ir) stdv_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
            xcov_3d[IX3(i, j, k, l, nx)] = ORC_AMISS;
          }
        }
  return 0;
}

/* GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341): the normalisation of a -tch (three-cornered hat) run, nr = 3 residual
 * series.  Means and rms as in GrITAS_Norm_, but the third output is the VARIANCE (no square root, :1274, :1322) and
 * xcov receives the 3CH estimates.  Kept literally: the whole-array 3CH expressions run over missing values too
 * (0.5*(AMISS+AMISS-AMISS)), the 2-D and the 3-D blocks combine the variances in a DIFFERENT order for slots 2 and 3
 * (:1287-1289 vs :1337-1339), a 3-D box with one observation keeps tmp (:1324-1325) while the 2-D one is missing, and
 * with nlevs < 2 the routine returns before the 3-D block (:1292). */
int orc_norm_3ch(int im, int jm, int km, int nlevs, int nr, int nrmlz, int l2d, double *bias_2d,
                 double *rtms_2d, double *var_2d, double *xcov_2d, double *nobs_2d, int l3d,
                 double *bias_3d, double *rtms_3d, double *var_3d, double *xcov_3d, double *nobs_3d,
                 int ntresh) {
  int i, j, k, l, m, ir;
  double tmp[3] = {0.0, 0.0, 0.0}, n;
  if (nr != 3) return 8;                      /* :1240 rejects nr > 3; the 3CH expressions need exactly three series */

  m = 1;                                                                   /* :1247 */
  for (l = 1; l <= l2d; ++l)
    for (j = 1; j <= jm; ++j)
      for (i = 1; i <= im; ++i) {
        n = nobs_2d[IX2(i, j, l, 1)];                                      /* :1252 */
        if (nrmlz) m = (int)n;                                             /* :1253 */
        if (n > 0.0 && n >= (double)ntresh) {                              /* :1254-1260 */
          for (ir = 1; ir <= nr; ++ir) {
            bias_2d[IX2(i, j, l, ir)] = bias_2d[IX2(i, j, l, ir)] / (double)m;
            tmp[ir - 1] = rtms_2d[IX2(i, j, l, ir)] / (double)m;
            rtms_2d[IX2(i, j, l, ir)] = sqrt(tmp[ir - 1]);
            tmp[ir - 1] = fabs(tmp[ir - 1] - bias_2d[IX2(i, j, l, ir)] * bias_2d[IX2(i, j, l, ir)]);
          }
        } else {                                                           /* :1261-1266 */
          for (ir = 1; ir <= nr; ++ir) {
            bias_2d[IX2(i, j, l, ir)] = ORC_AMISS;
            rtms_2d[IX2(i, j, l, ir)] = ORC_AMISS;
          }
        }
        if (n > 1.0 && n >= (double)ntresh) {                              /* :1269-1272 */
          for (ir = 1; ir <= nr; ++ir) var_2d[IX2(i, j, l, ir)] = (double)m * tmp[ir - 1] / (double)(m - 1);
        } else {                                                           /* :1273-1277 */
          for (ir = 1; ir <= nr; ++ir) var_2d[IX2(i, j, l, ir)] = ORC_AMISS;
        }
      }
  for (l = 1; l <= l2d; ++l)                                               /* :1287-1289 */
    for (j = 1; j <= jm; ++j)
      for (i = 1; i <= im; ++i) {
        const double v1 = var_2d[IX2(i, j, l, 1)], v2 = var_2d[IX2(i, j, l, 2)], v3 = var_2d[IX2(i, j, l, 3)];
        xcov_2d[IX2(i, j, l, 1)] = 0.5 * (v1 + v2 - v3);
        xcov_2d[IX2(i, j, l, 2)] = 0.5 * (v2 + v3 - v1);
        xcov_2d[IX2(i, j, l, 3)] = 0.5 * (v3 + v1 - v2);
      }

  if (nlevs < 2) return 0;                                                 /* :1292 */

  m = 1;                                                                   /* :1297 */
  for (l = 1; l <= l3d; ++l)
    for (k = 1; k <= nlevs; ++k)
      for (j = 1; j <= jm; ++j)
        for (i = 1; i <= im; ++i) {
          n = nobs_3d[IX3(i, j, k, l, 1)];                                 /* :1303 */
          if (nrmlz) m = (int)n;                                           /* :1304 */
          if (n > 0.0 && n >= (double)ntresh) {                            /* :1305-1311 */
            for (ir = 1; ir <= nr; ++ir) {
              bias_3d[IX3(i, j, k, l, ir)] = bias_3d[IX3(i, j, k, l, ir)] / (double)m;
              tmp[ir - 1] = rtms_3d[IX3(i, j, k, l, ir)] / (double)m;
              rtms_3d[IX3(i, j, k, l, ir)] = sqrt(tmp[ir - 1]);
              tmp[ir - 1] =
                  fabs(tmp[ir - 1] - bias_3d[IX3(i, j, k, l, ir)] * bias_3d[IX3(i, j, k, l, ir)]);
            }
          } else {                                                         /* :1312-1317 */
            for (ir = 1; ir <= nr; ++ir) {
              bias_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
              rtms_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
            }
          }
          if (n > 1.0 && n >= (double)ntresh) {                            /* :1319-1320 */
            for (ir = 1; ir <= nr; ++ir) var_3d[IX3(i, j, k, l, ir)] = (double)m * tmp[ir - 1] / (double)(m - 1);
          } else if (n == 1.0) {                                           /* :1322-1323: stale tmp if ntresh > 1 */
            for (ir = 1; ir <= nr; ++ir) var_3d[IX3(i, j, k, l, ir)] = tmp[ir - 1];
          } else {                                                         /* :1324-1325 */
            for (ir = 1; ir <= nr; ++ir) var_3d[IX3(i, j, k, l, ir)] = ORC_AMISS;
          }
        }
  for (l = 1; l <= l3d; ++l)                                               /* :1337-1339: ALL km levels */
    for (k = 1; k <= km; ++k)
      for (j = 1; j <= jm; ++j)
        for (i = 1; i <= im; ++i) {
          const double v1 = var_3d[IX3(i, j, k, l, 1)], v2 = var_3d[IX3(i, j, k, l, 2)], v3 = var_3d[IX3(i, j, k, l, 3)];
          xcov_3d[IX3(i, j, k, l, 1)] = 0.5 * (v1 + v2 - v3);
          xcov_3d[IX3(i, j, k, l, 2)] = 0.5 * (v3 + v1 - v2);
          xcov_3d[IX3(i, j, k, l, 3)] = 0.5 * (v2 + v3 - v1);
        }
  return 0;
}

/* gritas.f:1002-1003: nobs / nredundant (ensemble files valid at the same time). */
void orc_div_nredundant(size_t n, double *nobs_grid, int nredundant) {
  size_t i;
  for (i = 0; i < n; ++i) nobs_grid[i] = nobs_grid[i] / (double)nredundant;
}

/* m_gritas_grids.f:1019-1167 (GrITAS_MinMax_): nr must be 1 ("could not handle more then 1 field", :1072).
 * minv / maxv are inout (the driver starts them at AMISS / -AMISS, gritas.f:364-366).  Same return codes and
 * level-miss behaviour as orc_accum (the message at :1133-1134 is the Accum one). */
int orc_minmax(int im, int jm, int km, int nlevs, int nobs, int nr, const double *lat, const double *lon,
               const double *lev, const int *kx, const int *kt, const double *del, const int *kxkt_table, int kxmax,
               int ktmax, const double *p1, const double *p2, int *ob_flag, int l2d, double *minv_2d, double *maxv_2d,
               double *nobs_2d, int l3d, double *minv_3d, double *maxv_3d, double *nobs_3d, double *bad_lev, int *bad_n) {
  int n, i, j, k, kk, l, surface_var;
  double dlon, dlat;
  long ii, jj;
  if (nr > 1) return 8;                                                  /* :1071-1072 */
  orc_grid_init(im, jm, &dlon, &dlat);
  for (n = 0; n < nobs; ++n) {                                           /* :1077 */
    if (ob_flag[n] != 0) continue;                                       /* :1078-1079 */
    if (kx[n] < 1 || kx[n] > kxmax || kt[n] < 1 || kt[n] > ktmax) l = 0;
    else l = kxkt_table[(size_t)(kx[n] - 1) + (size_t)kxmax * (size_t)(kt[n] - 1)]; /* :1081 */
    surface_var = l < 0;
    l = abs(l);
    if (l <= 0) { ob_flag[n] = 1; continue; }                            /* :1085-1088 */
    ii = 1 + orc_nint((lon[n] - ORC_LONMIN) / dlon);                     /* :1092-1098 */
    if (ii == (long)im + 1) ii = 1;
    if (ii == 0) ii = im;
    jj = 1 + orc_nint((lat[n] - ORC_LATMIN) / dlat);
    if (ii < 1) ii = 1;
    if (ii > im) ii = im;
    if (jj < 1) jj = 1;
    if (jj > jm) jj = jm;
    i = (int)ii;
    j = (int)jj;
    if (surface_var) {                                                   /* :1104-1113 */
      if (del[n] < minv_2d[IX2(i, j, l, 1)]) minv_2d[IX2(i, j, l, 1)] = del[n];
      if (del[n] > maxv_2d[IX2(i, j, l, 1)]) maxv_2d[IX2(i, j, l, 1)] = del[n];
      nobs_2d[IX2(i, j, l, 1)] = nobs_2d[IX2(i, j, l, 1)] + 1.0;
    } else {
      k = 0;
      for (kk = 1; kk <= nlevs; ++kk)                                    /* :1123-1130 */
        if (lev[n] <= p1[kk - 1] && lev[n] > p2[kk - 1]) { k = kk; break; }
      if (k == 0) {                                                      /* :1132-1135 */
        if (bad_lev) *bad_lev = lev[n];
        if (bad_n) *bad_n = n;
        return 7;
      }
      if (del[n] < minv_3d[IX3(i, j, k, l, 1)]) minv_3d[IX3(i, j, k, l, 1)] = del[n];   /* :1143-1148 */
      if (del[n] > maxv_3d[IX3(i, j, k, l, 1)]) maxv_3d[IX3(i, j, k, l, 1)] = del[n];
      nobs_3d[IX3(i, j, k, l, 1)] = nobs_3d[IX3(i, j, k, l, 1)] + 1.0;
    }
  }
  return 0;
}

/* m_gritas_masks.F90:137-186 (maskout_): flags the observations whose box of the (im, jm) fraction field holds
 * less than 0.99.  Returns the number of observations excluded (the value the reference prints). */
int orc_maskout(int im, int jm, int nobs, const double *lat, const double *lon, const double *maskfld, int *qcx) {
  int n, nexcl = 0;
  double dlon, dlat;
  long ii, jj;
  orc_grid_init(im, jm, &dlon, &dlat);                                   /* :151-152 */
  for (n = 0; n < nobs; ++n) {
    ii = 1 + orc_nint((lon[n] - ORC_LONMIN) / dlon);                     /* :168-174 */
    if (ii == (long)im + 1) ii = 1;
    if (ii == 0) ii = im;
    jj = 1 + orc_nint((lat[n] - ORC_LATMIN) / dlat);
    if (ii < 1) ii = 1;
    if (ii > im) ii = im;
    if (jj < 1) jj = 1;
    if (jj > jm) jj = jm;
    if (maskfld[(size_t)(ii - 1) + (size_t)im * (size_t)(jj - 1)] < 0.99 && qcx[n] == 0) {   /* :178-181 */
      nexcl++;
      qcx[n] = 1;
    }
  }
  return nexcl;
}

/* ---------------------------------------------------------------------------------------------
 * Channel / level cross-covariance family (m_gritas_grids.f:667-1008, 1862-2059).  Arrays:
 *   bias_lv(ndim, l3d, nr), xcov_lv(ndim, ndim, l3d, nr), nobs_lv(ndim, l3d), column-major, 1-based. */
#define XB(k, l, r) bias_lv[((size_t)((k) - 1)) + (size_t)ndim * (((size_t)((l) - 1)) + (size_t)l3d * ((size_t)((r) - 1)))]
#define XC(a, b, l, r) \
  xcov_lv[((size_t)((a) - 1)) + (size_t)ndim * (((size_t)((b) - 1)) + (size_t)ndim * (((size_t)((l) - 1)) + (size_t)l3d * ((size_t)((r) - 1))))]
#define XN(k, l) nobs_lv[((size_t)((k) - 1)) + (size_t)ndim * ((size_t)((l) - 1))]

/* m_gritas_grids.f:667-848 (GrITAS_XCov_Accum_).  km = profile length (module variable), achan = the
 * achan(:, ichan_version) column of the active channels (m_gritas_grids.f:69-71), ndim = nchan.
 * Returns 0, or 1 'improper dim settings' (:734-736), 2 'TCH setting inconsistent w/ dims' (:738-740),
 * 3 'Desroziers setting inconsistent w/ dims' (:742-744), 4 'not all profiles complete' (:754-757),
 * 5 'error in channel matching' (:789-790).  lat / lon are arguments of the reference routine but never used. */
int orc_xcov_accum(int km, int nchan, const int *achan, int nobs, int nr, const double *lev, const int *kx,
                   const int *kt, const double *del, const int *kxkt_table, int kxmax, int ktmax, int *ob_flag,
                   int l3d, double *bias_lv, double *xcov_lv, double *nobs_lv, int *nprof, int tch) {
  const int ndim = nchan;
  int n1, ii, kk, k, m1, l, ir, rc = 0;
  int *indx, *ll, *reject_ob;
  double *rdel;
  (void)ktmax;
  if (nr > 3) return 1;
  if (tch) { if (nr != 3) return 2; } else { if (nr != 2) return 3; }
  if (nobs % km != 0) return 4;                                        /* :751-757 */
  indx = (int *)calloc((size_t)nchan, sizeof(int));
  ll = (int *)calloc((size_t)nchan, sizeof(int));
  reject_ob = (int *)calloc((size_t)nchan, sizeof(int));
  rdel = (double *)malloc(sizeof(double) * (size_t)nchan * (size_t)nr);
  *nprof = 0;
  for (n1 = 1; n1 <= nobs && !rc; n1 += km) {                          /* :774 */
    const int n2 = n1 + km - 1;
    int any_rej = 0;
    m1 = 0;
    for (kk = 0; kk < nchan; ++kk) reject_ob[kk] = 0;                  /* :777 */
    for (ii = n1; ii <= n2; ++ii) {                                    /* :778 */
      k = 0;
      for (kk = 1; kk <= nchan; ++kk)                                  /* :780-785: integer |achan - nint(lev)| < 1e-5 */
        if (labs((long)achan[kk - 1] - orc_nint(lev[ii - 1])) < 1e-5) { k = kk; break; }
      if (k > 0) {                                                     /* :787 */
        m1 = m1 + 1;
        if (m1 > nchan) { rc = 5; break; }                             /* :789-790 */
        indx[m1 - 1] = ii;
        reject_ob[m1 - 1] = ob_flag[ii - 1] != 0;
        ll[m1 - 1] = kxkt_table[(size_t)(kx[ii - 1] - 1) + (size_t)kxmax * (size_t)(kt[ii - 1] - 1)];
        ll[m1 - 1] = abs(ll[m1 - 1]);
        if (ll[m1 - 1] <= 0) {                                         /* :795-798 */
          ob_flag[ii - 1] = 1;
          reject_ob[m1 - 1] = 1;
        }
      }
    }
    if (rc) break;
    for (kk = 0; kk < nchan; ++kk) any_rej |= reject_ob[kk];
    if (any_rej) continue;                                             /* :801 */
    if (m1 != nchan) continue;                                         /* :802 */
    for (kk = 0; kk < nchan; ++kk)                                     /* :809 rdel(:,:) = del(indx(:),:) */
      for (ir = 0; ir < nr; ++ir) rdel[kk + (size_t)nchan * ir] = del[(size_t)(indx[kk] - 1) + (size_t)nobs * ir];
    *nprof = *nprof + 1;
    for (kk = 1; kk <= nchan; ++kk) {                                  /* :815 */
      l = ll[kk - 1];
      for (ir = 1; ir <= nr; ++ir) XB(kk, l, ir) = XB(kk, l, ir) + rdel[(kk - 1) + (size_t)nchan * (ir - 1)];
      if (tch) {                                                       /* :821-824 */
        for (ii = 1; ii <= nchan; ++ii)
          for (ir = 1; ir <= nr; ++ir)
            XC(ii, kk, l, ir) = XC(ii, kk, l, ir) + rdel[(ii - 1) + (size_t)nchan * (ir - 1)] * rdel[(kk - 1) + (size_t)nchan * (ir - 1)];
      } else {                                                         /* :826-830 */
        for (ii = 1; ii <= nchan; ++ii)
          XC(ii, kk, l, 1) = XC(ii, kk, l, 1) + rdel[(ii - 1)] * rdel[(kk - 1) + (size_t)nchan];
      }
      XN(kk, l) = XN(kk, l) + 1.0;                                     /* :832 */
    }
  }
  free(indx); free(ll); free(reject_ob); free(rdel);
  return rc;
}

/* m_gritas_grids.f:859-1008 (GrITAS_XCov_Norm_).  INTEGER m (:903); `m = n` truncates; m keeps its last value between
 * iterations when nrmlz is false (= 1: nothing is normalised).  Returns 0 or the gritas_perror cases 1..3 as above. */
int orc_xcov_norm(int ndim, int nr, int nrmlz, int tch, int self, int l3d, double *bias_lv, double *xcov_lv,
                  const double *nobs_lv, int ntresh) {
  int k1, k2, k3, l, m, i1, i2, nrr, ir;
  double xtmp = 0.0, n;
  if (nr > 3) return 1;
  if (tch) { if (nr != 3) return 2; } else { if (nr != 2) return 3; }
  nrr = tch ? 3 : 1;                                                   /* :936-937 */
  for (k3 = 1; k3 <= nrr; ++k3) {                                      /* :938-958 */
    m = 1;
    for (l = 1; l <= l3d; ++l)
      for (k2 = 1; k2 <= ndim; ++k2) {
        n = XN(k2, l);
        if (nrmlz) m = (int)n;
        if (m > 1 && m >= ntresh) {
          if (tch) XB(k2, l, k3) = XB(k2, l, k3) / m;
          else for (ir = 1; ir <= nr; ++ir) XB(k2, l, ir) = XB(k2, l, ir) / m;
        }
      }
    m = 1;
  }
  m = 1;
  for (k3 = 1; k3 <= nrr; ++k3) {                                      /* :959-989 */
    i1 = 1; i2 = 2;
    if (nrr == 3) { i1 = k3; i2 = k3; }
    for (l = 1; l <= l3d; ++l)
      for (k2 = 1; k2 <= ndim; ++k2)
        for (k1 = 1; k1 <= ndim; ++k1) {
          n = XN(k2, l);
          if (nrmlz) m = (int)n;
          if (m > 1 && m >= ntresh) {
            xtmp = XC(k1, k2, l, k3) / m;
            if (tch) xtmp = xtmp - XB(k1, l, i1) * XB(k2, l, i2);
            else xtmp = xtmp - 0.5 * (XB(k1, l, i1) * XB(k2, l, i2) + XB(k2, l, i1) * XB(k1, l, i2));
          }
          if (m > 1 && m >= ntresh) XC(k1, k2, l, k3) = m * xtmp / (m - 1);
        }
  }
  if (!tch) return 0;                                                  /* :991 */
  if (self) return 0;                                                  /* :992 */
  {                                                                    /* :994-1002 three-cornered hat */
    const size_t plane = (size_t)ndim * ndim * l3d;
    size_t e;
    for (e = 0; e < plane; ++e) {
      const double v1 = xcov_lv[e], v2 = xcov_lv[e + plane], v3 = xcov_lv[e + 2 * plane];
      xcov_lv[e] = 0.5 * (v1 + v2 - v3);
      xcov_lv[e + plane] = 0.5 * (v3 + v1 - v2);
      xcov_lv[e + 2 * plane] = 0.5 * (v2 + v3 - v1);
    }
  }
  return 0;
}

/* m_gritas_grids.f:1862-2059 (GrITAS_XCov_Prs_Accum_): profiles are runs of equal sounding index ks; every
 * sounding index met gathers ALL observations that carry it (:1973-1979), so an index that comes back later is
 * processed once per run.  ndim = nlevs, l = 1 (:2026).  kx / kt / kxkt_table / ob_flag are arguments of the reference
 * routine but never used.  (The reference sizes its scratch with a count that is off by one at run boundaries,
 * :1953-1963 -- undefined behaviour there; the restatement allocates what the gather needs.)
 * Returns 0, 1..3 as above, or 7 'could not find level' (:1990, :2012) with *bad_lev set. */
int orc_xcov_prs_accum(int nlevs, int nobs, int nr, const double *lev, const int *ks, const double *del,
                       const double *p1, const double *p2, int l3d, double *bias_lv, double *xcov_lv,
                       double *nobs_lv, int *nprof, int tch, double *bad_lev) {
  const int ndim = nlevs;
  int n, m, np, ip, kk, ks1, ir, *uks, *kindx;
  double *rdel;
  if (nr > 3) return 1;
  if (tch) { if (nr != 3) return 2; } else { if (nr != 2) return 3; }
  *nprof = 0;
  if (nobs <= 0) return 0;
  *nprof = 1;                                                          /* :1942-1949 */
  ks1 = ks[0];
  for (n = 1; n <= nobs; ++n)
    if (ks[n - 1] != ks1) { ks1 = ks[n - 1]; *nprof = *nprof + 1; }
  uks = (int *)malloc(sizeof(int) * (size_t)(*nprof));
  np = 1; ks1 = ks[0]; uks[0] = ks1;                                   /* :1952-1964 */
  for (n = 1; n <= nobs; ++n)
    if (ks[n - 1] != ks1) { ks1 = ks[n - 1]; np = np + 1; uks[np - 1] = ks1; }
  kindx = (int *)malloc(sizeof(int) * (size_t)nobs);
  rdel = (double *)malloc(sizeof(double) * (size_t)nobs * (size_t)nr);
  for (np = 1; np <= *nprof; ++np) {                                   /* :1974 */
    ip = 0;
    for (n = 1; n <= nobs; ++n) {                                      /* :1977-1982 */
      if (ks[n - 1] != uks[np - 1]) continue;
      ip = ip + 1;
      kindx[ip - 1] = 0;
      for (kk = 1; kk <= nlevs; ++kk)                                  /* :2004-2013 */
        if (lev[n - 1] <= p1[kk - 1] && lev[n - 1] > p2[kk - 1]) { kindx[ip - 1] = kk; break; }
      if (kindx[ip - 1] == 0) {
        if (bad_lev) *bad_lev = lev[n - 1];
        free(uks); free(kindx); free(rdel);
        return 7;
      }
      for (ir = 0; ir < nr; ++ir) rdel[(ip - 1) + (size_t)nobs * ir] = del[(n - 1) + (size_t)nobs * ir];
    }
    {
      const int l = 1;                                                 /* :2026 */
      for (n = 1; n <= ip; ++n) {                                      /* :2027 */
        const int k1 = kindx[n - 1];
        XN(k1, l) = XN(k1, l) + 1.0;                                   /* :2034 */
        for (ir = 1; ir <= nr; ++ir) XB(k1, l, ir) = XB(k1, l, ir) + rdel[(n - 1) + (size_t)nobs * (ir - 1)];
        for (m = 1; m <= ip; ++m) {                                    /* :2036 */
          const int k2 = kindx[m - 1];
          if (tch) {
            for (ir = 1; ir <= nr; ++ir)
              XC(k1, k2, l, ir) = XC(k1, k2, l, ir) + rdel[(n - 1) + (size_t)nobs * (ir - 1)] * rdel[(m - 1) + (size_t)nobs * (ir - 1)];
          } else {
            XC(k1, k2, l, 1) = XC(k1, k2, l, 1) + 0.5 * (rdel[(n - 1)] * rdel[(m - 1) + (size_t)nobs] + rdel[(n - 1) + (size_t)nobs] * rdel[(m - 1)]);
          }
        }
      }
    }
  }
  free(uks); free(kindx); free(rdel);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Scorecard reductions.  m_gritas_grids.f:163-175 (init_): first / last latitude row of a region. */
void orc_score_regions(int jm, int nregs, const double *regbounds, int *regindx) {
  const double dlat = 180.0 / (double)(jm - 1);
  int i, j;
  for (i = 0; i < nregs; ++i) {
    regindx[2 * i] = 0; regindx[2 * i + 1] = 0;
    for (j = 1; j <= jm; ++j) {
      const double glat = ORC_LATMIN + (j - 1) * dlat;              /* :136-138 */
      if (glat <= regbounds[2 * i]) regindx[2 * i] = j;
      if (glat > 0.0 && glat <= regbounds[2 * i + 1]) regindx[2 * i + 1] = j;
      if (glat < 0.0 && glat <= regbounds[2 * i + 1]) regindx[2 * i + 1] = j;
    }
  }
}

/* m_gritas_grids.f:1797-1826 (ave_nomiss): mean of x over the box range where nobs >= 0 and x is not AMISS. */
static double orc_ave_nomiss(const double *x, const double *nobs, int im, const int *xregion, const int *yregion) {
  const double threshold = 0.0;
  double add = -ORC_AMISS, counter = 0.0;
  int ii, jj;
  for (jj = yregion[0]; jj <= yregion[1]; ++jj)
    for (ii = xregion[0]; ii <= xregion[1]; ++ii) {
      const size_t o = (size_t)(ii - 1) + (size_t)im * (size_t)(jj - 1);
      if (nobs[o] >= threshold && fabs(x[o] - ORC_AMISS) > 0.01) {
        if (add == -ORC_AMISS) add = x[o]; else add = add + x[o];
        counter = counter + 1.0;
      }
    }
  return add == -ORC_AMISS ? ORC_AMISS : add / counter;
}

/* m_gritas_grids.f:1671-1795 (scores2_), nr = 1, nt = 1: collect_stats(nregs, km, nv2d + nv3d, nevals) real(4);
 * type_stats = mean, rms, stdv (:64); surface variables first, level slot 1 only. */
void orc_scores2(int im, int jm, int km, int nv2d, int nv3d, int nregs, const int *regindx, const int *lonindx,
                 const double *bias_2d, const double *rtms_2d, const double *stdv_2d, const double *nobs_2d,
                 const double *bias_3d, const double *rtms_3d, const double *stdv_3d, const double *nobs_3d,
                 float *collect_stats) {
  const size_t plane = (size_t)im * jm;
  const int nvar = nv2d + nv3d;
  int ne, mr, lm, kk, nv;
  for (ne = 1; ne <= 3; ++ne) {
    const double *s2 = ne == 1 ? bias_2d : (ne == 2 ? rtms_2d : stdv_2d);
    const double *s3 = ne == 1 ? bias_3d : (ne == 2 ? rtms_3d : stdv_3d);
    for (mr = 1; mr <= nregs; ++mr) {
      nv = 0;
      for (lm = 1; lm <= nv2d; ++lm) {
        nv = nv + 1;
        collect_stats[(mr - 1) + (size_t)nregs * (0 + (size_t)km * ((nv - 1) + (size_t)nvar * (ne - 1)))] =
            (float)orc_ave_nomiss(s2 + plane * (lm - 1), nobs_2d + plane * (lm - 1), im, lonindx, regindx + 2 * (mr - 1));
      }
      for (lm = 1; lm <= nv3d; ++lm) {
        nv = nv + 1;
        for (kk = 1; kk <= km; ++kk)
          collect_stats[(mr - 1) + (size_t)nregs * ((kk - 1) + (size_t)km * ((nv - 1) + (size_t)nvar * (ne - 1)))] =
              (float)orc_ave_nomiss(s3 + plane * ((kk - 1) + (size_t)km * (lm - 1)), nobs_3d + plane * ((kk - 1) + (size_t)km * (lm - 1)),
                                    im, lonindx, regindx + 2 * (mr - 1));
      }
    }
  }
}
