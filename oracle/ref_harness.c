/* Glue around oracle/_ref/gritas_ref_gen.c -- the C that oracle/f2c_subset.py generates from the reference's own
 * Fortran source (m_gritas_grids.f), statement by statement.  TEST INFRASTRUCTURE: it pins the hand-written oracle
 * (tests/test_oracle_vs_ref.py) and the golden vectors (tests/golden/make_golden.py); nothing under gritas_b200/ uses it.
 *
 * Only what a Fortran caller supplies is written here: the module state grids_init / the rc file set up
 * (m_gritas_grids.f:105-190: im, jm, km, nlevs, plevs; KXMAX / KTMAX of m_odsmeta), the optional argument, and
 * gritas_perror (etc.f:54-60: print and exit(7)) turned into a return code of 7. */
#include "_ref/gritas_ref_gen.c"
#include <string.h>

int ref_init(int im_, int jm_, int km_, int nlevs_, const double *plevs_, int kxmax_, int ktmax_) {
  im = im_; jm = jm_; km = km_; nlevs = nlevs_; kxmax = kxmax_; ktmax = ktmax_;
  free(plevs);
  plevs = (double *)malloc(sizeof(double) * (size_t)(km_ > 0 ? km_ : 1));      /* allocate ( ..., plevs(km), ... ) :120 */
  memcpy(plevs, plevs_, sizeof(double) * (size_t)nlevs_);
  gritas_ref_mesh();                                                            /* dLon, dLat :128-129 */
  return 0;
}
double ref_dlon(void) { return dlon; }
double ref_dlat(void) { return dlat; }
int ref_error_line(void) { return gritas_ref_error_line; }

#define GUARD() do { gritas_ref_error_line = 0; if (setjmp(gritas_ref_jmp)) return 7; } while (0)

int ref_pint(double *p1, double *p2) { GUARD(); gritas_pint_(p1, p2); return 0; }

int ref_accum(double *lat, double *lon, double *lev, int *kx, int *kt, double *del, int nobs, int nr, int *kxkt_table,
              double *p1, double *p2, int *ob_flag, int l2d, double *bias_2d, double *rtms_2d, double *xcov_2d,
              double *nobs_2d, int l3d, double *bias_3d, double *rtms_3d, double *xcov_3d, double *nobs_3d) {
  GUARD();
  gritas_accum_(lat, lon, lev, kx, kt, del, nobs, nr, kxkt_table, p1, p2, ob_flag, l2d, bias_2d, rtms_2d, xcov_2d, nobs_2d,
                l3d, bias_3d, rtms_3d, xcov_3d, nobs_3d);
  return 0;
}

int ref_minmax(double *lat, double *lon, double *lev, int *kx, int *kt, double *del, int nobs, int nr, int *kxkt_table,
               double *p1, double *p2, int *ob_flag, int l2d, double *minv_2d, double *maxv_2d, double *nobs_2d, int l3d,
               double *minv_3d, double *maxv_3d, double *nobs_3d) {
  GUARD();
  gritas_minmax_(lat, lon, lev, kx, kt, del, nobs, nr, kxkt_table, p1, p2, ob_flag, l2d, minv_2d, maxv_2d, nobs_2d, l3d,
                 minv_3d, maxv_3d, nobs_3d);
  return 0;
}

/* has_ntresh = 0: the optional argument is absent (gritas.f:885-887 calls GrITAS_Norm without it) */
int ref_norm(int nr, int nrmlz, int l2d, double *bias_2d, double *rtms_2d, double *stdv_2d, double *xcov_2d, double *nobs_2d,
             int l3d, double *bias_3d, double *rtms_3d, double *stdv_3d, double *xcov_3d, double *nobs_3d, int ntresh,
             int has_ntresh) {
  GUARD();
  gritas_norm_(nr, nrmlz, l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d, l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d,
               has_ntresh ? &ntresh : (int *)0);
  return 0;
}

int ref_norm_3ch(int nr, int nrmlz, int l2d, double *bias_2d, double *rtms_2d, double *var_2d, double *xcov_2d,
                 double *nobs_2d, int l3d, double *bias_3d, double *rtms_3d, double *var_3d, double *xcov_3d, double *nobs_3d,
                 int ntresh, int has_ntresh) {
  GUARD();
  gritas_3ch_norm_(nr, nrmlz, l2d, bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d, l3d, bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d,
                   has_ntresh ? &ntresh : (int *)0);
  return 0;
}

/* is_surface (gritas_kxkt.f:121-171): kt < 0 or kt among ktSurfAll of m_odsmeta (GMAO_ods, not in the tree: the list is
 * an input of the harness, as it is of the oracle) */
static int n_surf_ = 0, kt_surf_[64];
void ref_set_surface_kts(const int *kts, int n) { n_surf_ = n > 64 ? 64 : n; memcpy(kt_surf_, kts, sizeof(int) * (size_t)n_surf_); }
int is_surface_(int kt) {
  int ic = 0;
  for (int i = 0; i < n_surf_; ++i) ic += kt_surf_[i] == kt;   /* ic(1) = count(ktSurfAll.eq.kt) */
  return ic > 0 || kt < 0;
}

int ref_kxkt(int ktmax_, int kxmax_, int l2d_max, int l3d_max, int listsz, int *kt_list, int *kx_list, int *kxkt_table,
             int *list_2d, int *list_3d, int *l2d, int *l3d) {
  GUARD();
  gritas_kxkt_(ktmax_, kxmax_, l2d_max, l3d_max, listsz, kt_list, kx_list, kxkt_table, list_2d, list_3d, l2d, l3d);
  return 0;
}

int ref_select(int iopt, int scale_res, int xcov, int pxcov, int tch, int self, int meanres, int nobs, int nr, int x_passive,
               int *nstat, double *del, double *omf, double *oma, double *obs, double *xvec, double *xm, int *qcexcl,
               double *m_obs, double *m_omf, double *m_oma) {
  GUARD();
  gritas_select_(iopt, scale_res, xcov, pxcov, tch, self, meanres, nobs, nr, x_passive, nstat, del, omf, oma, obs, xvec, xm,
                 qcexcl, m_obs, m_omf, m_oma);
  return 0;
}

int ref_filter(int nobs, int ioptn, int t1, int t2, double lona, double lonb, int x_passive, int x_pre_bad, int *qcexcl,
               int *time, double *lon, int *ob_flag) {
  GUARD();
  gritas_filter_(nobs, ioptn, t1, t2, lona, lonb, x_passive, x_pre_bad, qcexcl, time, lon, ob_flag);
  return 0;
}

int ref_jo_scale(int iopt, int nobs, int nr, double *del) { GUARD(); gritas_jo_scale_(iopt, nobs, nr, del); return 0; }

/* ---- channel / level cross-covariances (m_gritas_grids.f:667-848, 859-1008, 1862-2059) ----
 * active channels: achan(nchan, 2) = {index, number} columns and ichan_version pick one (set up by grids_init from the rc
 * file); the harness fills both columns with the same list, as the oracle's single list does */
void ref_set_channels(const int *ach, int n, int version) {
  free(achan);
  nchan = n; achan_n1 = n; ichan_version = version;
  achan = (int *)malloc(sizeof(int) * (size_t)(2 * (n > 0 ? n : 1)));
  for (int i = 0; i < n; ++i) { achan[i] = ach[i]; achan[n + i] = ach[i]; }
}

int ref_xcov_accum(double *lat, double *lon, double *lev, int *kx, int *kt, double *del, int nobs, int nr, int *kxkt_table,
                   int *ob_flag, int l3d, double *bias_lv, int b1, int b2, int b3, double *xcov_lv, double *nobs_lv, int *nprof,
                   int tch) {
  GUARD();
  gritas_xcov_accum_(lat, lon, lev, kx, kt, del, nobs, nr, kxkt_table, ob_flag, l3d, bias_lv, b1, b2, b3, xcov_lv, nobs_lv, nprof, tch);
  return 0;
}

int ref_xcov_norm(int nr, int nrmlz, int tch, int self, int l3d, double *bias_3d, int b1, int b2, int b3, double *xcov_3d, int x1,
                  int x2, int x3, int x4, double *nobs_3d, int n1, int n2, int ntresh, int has_ntresh) {
  GUARD();
  gritas_xcov_norm_(nr, nrmlz, tch, self, l3d, bias_3d, b1, b2, b3, xcov_3d, x1, x2, x3, x4, nobs_3d, n1, n2,
                    has_ntresh ? &ntresh : (int *)0);
  return 0;
}

int ref_xcov_prs_accum(double *lev, int *kx, int *kt, int *ks, double *del, int nobs, int nr, int *kxkt_table, double *p1,
                       double *p2, int *ob_flag, int l3d, double *bias_lv, int b1, int b2, int b3, double *xcov_lv, int x1, int x2,
                       int x3, int x4, double *nobs_lv, int n1, int n2, int *nprof, int tch) {
  GUARD();
  gritas_xcov_prs_accum_(lev, kx, kt, ks, del, nobs, nr, kxkt_table, p1, p2, ob_flag, l3d, bias_lv, b1, b2, b3, xcov_lv, x1, x2, x3,
                         x4, nobs_lv, n1, n2, nprof, tch);
  return 0;
}

/* ---- scorecard: ave_nomiss (m_gritas_grids.f:1797-1826) and the latitude rows of the regions (init_, :136-138, 163-175) ---- */
double ref_ave_nomiss(double *x, double *nobs, int n1, int n2, int *xregion, int *yregion) {
  return gritas_ave_nomiss_(x, n1, n2, nobs, n1, n2, xregion, yregion);
}

/* regbounds (2, nreg): what init_ wires for its region names (:146-157); regindx (2, nreg) starts at zero */
void ref_region_rows(int nreg, double *regbounds, int *regindx) {
  double *glat = (double *)malloc(sizeof(double) * (size_t)(jm > 0 ? jm : 1));
  gritas_glat_(glat);
  for (int i = 1; i <= nreg; ++i) gritas_region_rows_(i, nreg, glat, regbounds, regindx);   /* do i=1,size(regnames) :162 */
  free(glat);
}

/* ---- land / ocean mask gather: maskout_ (m_gritas_masks.F90:137-186) on a mask field the caller holds ---- */
int ref_maskout(int im_, int jm_, int nobs, double *lat, double *lon, double *maskfld, int *qcx) {
  double dlon_ = 0, dlat_ = 0;
  int nexcl = 0;                                     /* nexcl=0 :148 */
  gritas_mask_mesh_(im_, jm_, &dlon_, &dlat_);      /* dLon, dLat :152-153 */
  gritas_mask_loop_(nobs, im_, jm_, dlon_, dlat_, lat, lon, qcx, maskfld, &nexcl);
  return nexcl;
}

/* ---- the pre-sort (gritas.f:500-508).  IndexSet / IndexSort are m_MergeSorts of GMAO_mpeu (not in the tree): a STABLE merge
 * sort of the index by the key is ASSUMED here as in the oracle; what the translation pins is the sequence of keys, their
 * arrays and the direction.  is() is 1-based, as the Fortran index. ---- */
void indexset_(int n, int *is) { for (int i = 0; i < n; ++i) is[i] = i + 1; }
static void merge_sort_idx(int n, int *is, const double *key, int descend) {
  int *tmp = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b = mid, o = lo;
      while (a < mid && b < hi) {
        const double ka = key[is[a] - 1], kb = key[is[b] - 1];
        const int take_b = descend ? (kb > ka) : (kb < ka);               /* ties keep the earlier entry: stable */
        tmp[o++] = take_b ? is[b++] : is[a++];
      }
      while (a < mid) tmp[o++] = is[a++];
      while (b < hi) tmp[o++] = is[b++];
    }
    memcpy(is, tmp, sizeof(int) * (size_t)n);
  }
  free(tmp);
}
void indexsort_d_(int n, int *is, double *key, int descend) { merge_sort_idx(n, is, key, descend); }
void indexsort_i_(int n, int *is, int *key, int descend) {
  double *k = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) k[i] = (double)key[i];
  merge_sort_idx(n, is, k, descend);
  free(k);
}
int ref_presort(int nobs, int *is, int *qcexcl, double *lat, double *lon, double *lev, int *time, int *kt, int *ks, int *kx) {
  gritas_presort_(nobs, is, qcexcl, lat, lon, lev, time, kt, ks, kx);
  return 0;
}

/* scores2_ (m_gritas_grids.f:1671-1795).  Module state it reads: lonindx(2, ntslots) = (1, im) as init_ leaves it (:181-186)
 * and regindx(2, nreg) (the caller's, e.g. from ref_region_rows).  Shapes: 2-D arrays (im, jm, l2d, nr), counts (im, jm, l2d);
 * 3-D arrays (im, jm, km, l3d, nr), counts (im, jm, km, l3d); collect_stats (nreg, km, l2d + l3d, 3) real(4). */
int ref_scores2(float *collect_stats, int nreg, int im_, int jm_, int km_, int l2d, int l3d, int nr, int *regindx_,
                double *bias_2d, double *rtms_2d, double *stdv_2d, double *nobs_2d, double *bias_3d, double *rtms_3d,
                double *stdv_3d, double *nobs_3d) {
  static int lon_[2];
  lon_[0] = 1; lon_[1] = im_;
  lonindx = lon_; lonindx_n2 = 1;
  regindx = regindx_; regindx_n2 = nreg;
  GUARD();
  gritas_scores2_(collect_stats, nreg, km_, l2d + l3d, 3, bias_2d, im_, jm_, l2d, nr, rtms_2d, im_, jm_, l2d, nr, stdv_2d, im_, jm_, l2d, nr,
                  nobs_2d, im_, jm_, l2d, bias_3d, im_, jm_, km_, l3d, nr, rtms_3d, im_, jm_, km_, l3d, nr, stdv_3d, im_, jm_, km_, l3d, nr,
                  nobs_3d, im_, jm_, km_, l3d);
  return 0;
}
