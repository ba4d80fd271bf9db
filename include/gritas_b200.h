/*
 * gritas_b200.h -- C ABI of the B200-native GrITAS gridding / accumulation / finalize path.
 *
 * This is the drop-in boundary.  The reference (GEOS-ESM/GrITAS) has no FFI; its seam is the
 * pair of module generics GrITAS_Accum / GrITAS_Norm of m_gritas_grids
 * (src/Components/gritas/m_gritas_grids.f:88-89), called from
 * src/Applications/GrITAS_App/gritas.f:777-781 and 885-887 (and grisas.f:619-623, 639-642).
 * A replacement m_gritas_grids keeps those Fortran signatures and forwards to the entry points
 * below through ISO_C_BINDING (fortran/m_gritas_b200.F90).
 *
 * All arrays are caller-owned, contiguous, column-major and 1-based on the Fortran side.  Every
 * `real` of the reference is 64-bit (FREAL8, src/Components/gritas/CMakeLists.txt:3-4), every
 * `integer` 32-bit.  Input pointers may be pageable host, pinned host or device memory; the
 * library detects which and stages accordingly.  No entry point ever exits the process or falls
 * back to a CPU implementation: without a usable CUDA device gritas_b200_create fails with
 * GRITAS_B200_ERR_CUDA.
 */
#ifndef GRITAS_B200_H
#define GRITAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gritas_b200_ctx *gritas_b200_handle;

/* accumulation mode (low byte of `mode`) */
#define GRITAS_B200_MODE_FAST 0          /* summation order not fixed (<= 1e-6, measured ~1e-15): tiled accumulate (tile lists +
                                            shared-memory histograms) or warp-aggregated fp64 atomics, chosen from the data */
#define GRITAS_B200_MODE_DETERMINISTIC 1 /* every box is summed in input order (call order, then index within the call):
                                            bit-reproducible, the same sequence of fp64 additions as the reference's
                                            sequential loop.  Tiled path: the parked entries carry sequence numbers and the
                                            tile kernel orders every box's run by them; GRITAS_B200_FLAG_DET_SORT selects
                                            the older whole-file radix sort + in-order segment sums instead */
/* flags OR-ed into `mode` */
#define GRITAS_B200_FLAG_ASYNC 0x100     /* accum returns once inputs are consumed/staged; a level miss is
                                            reported by the next accum/sync/norm instead */
#define GRITAS_B200_FLAG_NO_AGGREGATE 0x200 /* fast mode: skip the __match_any_sync warp aggregation */
#define GRITAS_B200_FLAG_NO_STAGING 0x800   /* fast mode: never use the tiled accumulate (no tile lists are allocated); every
                                            observation goes to the HBM records with fp64 RED */
#define GRITAS_B200_FLAG_FORCE_TILED 0x1000 /* fast mode: use the tiled accumulate from the first call on instead of
                                            deciding from the locality of the first large call */
#define GRITAS_B200_FLAG_MINMAX 0x2000      /* the handle implements GrITAS_MinMax_ (m_gritas_grids.f:1019-1167, -minmax):
                                            accum keeps min / max / count per box instead of sums; nr must be 1 */
#define GRITAS_B200_FLAG_NO_WRITEBACK 0x4000 /* gritas_b200_accum reads ob_flag but does not write it back: only GrITAS_LBTally
                                            (-lb, gritas.f:791-796) consumes the flags GrITAS_Accum_ sets (m_gritas_grids.f:413-416);
                                            without it the download and, with ASYNC, the synchronisation per call go away */
#define GRITAS_B200_FLAG_DET_SORT 0x8000     /* deterministic mode: sort every call by box key (stable radix sort) and add the runs in
                                            order to the records, instead of the tiled path */
#define GRITAS_B200_FLAG_HOLD_INPUTS 0x400  /* with ASYNC: the caller keeps pinned input arrays valid until the
                                            next sync/norm, so accum does not even wait for the H2D copies */

/* return codes */
#define GRITAS_B200_OK 0
#define GRITAS_B200_ERR_LEVEL 7   /* 'Accum: could not find level' -- the reference exits with 7
                                     (m_gritas_grids.f:458-459 -> gritas_etc.f:54-60) */
#define GRITAS_B200_ERR_NR 8      /* 'could not handle more then 2 residuals' (m_gritas_grids.f:395, 555) */
#define GRITAS_B200_ERR_PROFILE 9 /* 'XCov_Accum: not all profiles complete' / 'error in channel matching'
                                     (m_gritas_grids.f:754-757, 789-790) */
#define GRITAS_B200_ERR_ORDER 10  /* deterministic mode, tiled path: a tile list overflowed, or one box held more parked
                                     observations than the tile kernel can order at once -- the statistics are complete but
                                     those observations were added out of input order (not reproducible) */
#define GRITAS_B200_ERR_CUDA (-1) /* CUDA failure or no device; see gritas_b200_last_error */
#define GRITAS_B200_ERR_ARG (-2)  /* bad argument */
#define GRITAS_B200_ERR_NCCL (-3) /* NCCL not loadable / NCCL failure */

/* Residual selection of the fused ODS entry (gritas.f:562-587).  |iopt| as in gritas.f:1408-1481. */
#define GRITAS_B200_IOPT_OMF 1
#define GRITAS_B200_IOPT_OMA 3
#define GRITAS_B200_IOPT_OMF_OMA 101 /* nr=2: del(:,1)=omf, del(:,2)=oma -- the reference's two runs
                                        (-omf, -oma; GrITAS_Proc/build-yaml.py.in:58) in one pass */

/*
 * Replaces: module state of m_gritas_grids (m_gritas_grids.f:24-72, init_ :105-190), the grid
 * allocation + zeroing of gritas.f:336-363, and the table arguments of GrITAS_Accum_
 * (m_gritas_grids.f:307-311).  dLon = 360/im, dLat = 180/(jm-1), LonMin = -180, LatMin = -90 are
 * derived exactly as m_gritas_grids.f:128-129, 33, 42.
 *   km          allocated vertical dimension of the 3-D grids (gritas.f:1885 sets km = nlevs)
 *   p1,p2       level intervals from GrITAS_Pint (m_gritas_grids.f:217-296), nlevs entries
 *   kxkt_table  (kxmax, ktmax) column-major: +l upper-air grid l, -l surface grid l, 0 unwanted
 *               (gritas_kxkt.f:105)
 *   nr          number of residual columns, 1..3 (3 = three-corner hat, xcov untouched)
 *   device      CUDA device ordinal, or -1 for LOCAL_RANK (env) / current device
 */
int gritas_b200_create(gritas_b200_handle *h, int im, int jm, int km, int nlevs, int l2d, int l3d, int nr,
                       const double *p1, const double *p2, const int32_t *kxkt_table, int kxmax, int ktmax,
                       int mode, int device);

/*
 * Host-side table builders of the path (pure host code, no device needed), for callers that do not
 * keep the Fortran originals:
 *   gritas_b200_pint  = GrITAS_Pint_ (m_gritas_grids.f:217-296): plevs (descending, hPa) -> p1, p2.
 *       returns 0, or 1 'must have at least 1 vertical levels', 2 'pressure levels do not decrease
 *       monotonically', 3 'cannot handle 0 hPa' (the three gritas_perror cases).
 *   gritas_b200_kxkt  = GrITAS_KxKt (gritas_kxkt.f:10-119): kt_list[listsz], kx_list (kxmax, listsz)
 *       column-major with a negative terminator per column, kt_surf[n_surf] = m_odsmeta::ktSurfAll ->
 *       kxkt_table (kxmax, ktmax), list_2d, list_3d, l2d, l3d.  returns 0, or 1 kt too large, 2 kt < 1,
 *       3 l2d too large, 4 l3d too large, 5 kx too large.
 */
int gritas_b200_pint(int nlevs, const double *plevs, double *p1, double *p2);
int gritas_b200_kxkt(int ktmax, int kxmax, int l2d_max, int l3d_max, int listsz, const int32_t *kt_list,
                     const int32_t *kx_list, const int32_t *kt_surf, int n_surf, int32_t *kxkt_table,
                     int32_t *list_2d, int32_t *list_3d, int *l2d, int *l3d);

/* Replaces GrITAS_Clean's deallocation of the grids (gritas.f:2017-2026). */
void gritas_b200_destroy(gritas_b200_handle h);

/*
 * Replaces GrITAS_Accum_ (m_gritas_grids.f:307-478), same argument meaning:
 *   del      (nobs, nr) column-major, leading dimension nobs
 *   ob_flag  inout, may be NULL (= all zero on input, nothing written back): nonzero on input
 *            rejects the observation; set to 1 on output when (kx,kt) is not in the table
 *   bad_lev  out, may be NULL: the offending level on GRITAS_B200_ERR_LEVEL (what the reference
 *            prints in 'Accum: obs level is ... hPa')
 * The grids live on the device and are only written to host arrays by _norm / _get_raw.
 */
int gritas_b200_accum(gritas_b200_handle h, int nobs, const double *lat, const double *lon, const double *lev,
                      const double *del, const int32_t *kx, const int32_t *kt, int32_t *ob_flag, double *bad_lev);

/*
 * Fused entry: residual selection (gritas.f:562-587, iopt above, scale_res 0..3), the QC / time /
 * longitude filter (gritas.f:710-739) and GrITAS_Accum_ in one kernel, straight from the ODS
 * columns.  time may be NULL when t2 < t1 (no time window); obs/xvec may be NULL when
 * scale_res == 0; oma may be NULL for IOPT_OMF.  ioptn > 0 keeps passive data (gritas.f:713-715).
 * x_passive / x_pre_bad are m_odsmeta constants (not in the reference tree), passed by the caller.
 * ob_flag_out (may be NULL) receives the flags GrITAS_Accum_ would leave behind.
 */
int gritas_b200_accum_ods(gritas_b200_handle h, int nobs, const double *lat, const double *lon, const double *lev,
                          const int32_t *time, const int32_t *kt, const int32_t *kx, const int32_t *qcexcl,
                          const double *omf, const double *oma, const double *obs, const double *xvec,
                          int iopt, int scale_res, int ioptn, int t1, int t2, double lona, double lonb,
                          int x_passive, int x_pre_bad, int32_t *ob_flag_out, double *bad_lev);

/*
 * Fused entry for all residual options of gritas.f:562-704 that work on one ODS structure: |iopt| as in the
 * reference (0/1 omf, 2/3 oma, 4/5 obs, 6/7 sigo, 8/9 obias, 10/11 hbh, 12/13 hah, 14/15 esigo, 16-19 job/joa,
 * 20/21 obs+xm, 22/23 imp, 24/25 obs-omf, 28/29 obs-omf+xm, 30/31 obs-oma, 32/33 (omf-oma)/xvec, oma/xvec) or
 * GRITAS_B200_IOPT_OMF_OMA.  The expressions are evaluated on the device exactly as written in the reference
 * (left to right, fp64).  options: GRITAS_B200_OPT_TCH (-tch, three residual columns for 10..15),
 * GRITAS_B200_OPT_SELF.  The -meanres variants (24..27, 30/31 with a second ODS) are not covered: build del on the
 * host and call gritas_b200_accum.  Columns an option does not use may be NULL.  The handle's nr must match.
 */
typedef struct gritas_b200_ods {
  const double *lat, *lon, *lev;
  const double *obs, *omf, *oma, *xvec, *xm;
  const int32_t *time, *kt, *kx, *qcexcl;
} gritas_b200_ods;
#define GRITAS_B200_OPT_TCH 1
#define GRITAS_B200_OPT_SELF 2
int gritas_b200_accum_odsx(gritas_b200_handle h, int nobs, const gritas_b200_ods *ods, int iopt, int options, int scale_res,
                           int ioptn, int t1, int t2, double lona, double lonb, int x_passive, int x_pre_bad,
                           int32_t *ob_flag_out, double *bad_lev);

/*
 * The -meanres variants (gritas.f:653-677, 680-691: a member ODS against the ensemble-mean ODS in the same order):
 * |iopt| 24/25 del = <omf> - omf; 26/27 del(:,1) = <omf> - omf, del(:,2) = xvec (nr = 2); 30/31 del = <oma> - oma.
 * mean_obs / mean_omf / mean_oma are the mean file's columns.  Where obs differs between the two files the observation is
 * made passive before the filter, as the reference does (the reference leaves del unset there; here it is the same
 * difference -- it only matters when passive data are kept).
 */
int gritas_b200_accum_odsx_meanres(gritas_b200_handle h, int nobs, const gritas_b200_ods *ods, const double *mean_obs,
                                   const double *mean_omf, const double *mean_oma, int iopt, int ioptn, int t1, int t2,
                                   double lona, double lonb, int x_passive, int x_pre_bad, int32_t *ob_flag_out,
                                   double *bad_lev);

/* Wait for all queued work; returns a pending GRITAS_B200_ERR_LEVEL (and its level) if any. */
int gritas_b200_sync(gritas_b200_handle h, double *bad_lev);

/* Same wait, but the tile lists of the tiled fast path stay parked (nothing is flushed into the records): what a host
 * that drives several handles needs before a sharded Norm, and what an ASYNC caller uses to pick up a level miss. */
int gritas_b200_wait(gritas_b200_handle h, double *bad_lev);

/* Deterministic mode: the place of the NEXT accumulate call in the summation order (default: calls since the last
 * reset).  A multi-rank run that shards one month f mod nranks passes the global file index, so that the sharded Norm adds
 * every box's observations in the order of the single-process file loop (gritas.f:389-393) whatever the rank count. */
int gritas_b200_set_call_order(gritas_b200_handle h, uint64_t order);

/* Tiled fast path only: accumulate every staged observation into the HBM records now (asynchronous;
 * sync / norm / get_raw / allreduce do it implicitly).  A no-op in the other modes. */
int gritas_b200_flush(gritas_b200_handle h);

/* The compute stream (gritas_b200_stream) waits for all accumulate work queued so far; no host
 * synchronisation.  The tiled fast path runs consecutive Accum calls on internal worker streams. */
int gritas_b200_barrier(gritas_b200_handle h);

/* Zero the device grids (grisas.f:285-296 zeroes them every synoptic time). */
int gritas_b200_reset(gritas_b200_handle h);

/*
 * Replaces GrITAS_Norm_ (m_gritas_grids.f:490-656): finalizes on the device and writes the
 * caller's arrays: bias (mean), rtms (root mean square), stdv, xcov(...,1) = <xy>,
 * xcov(...,nx) = covariance, nobs.  Shapes: *_2d (im,jm,l2d,nr), nobs_2d (im,jm,l2d),
 * *_3d (im,jm,km,l3d,nr), nobs_3d (im,jm,km,l3d).  Any pointer may be NULL (not wanted).
 * ntresh < 0 means "absent" (gritas.f:885-887 passes none).  In a multi-rank job call
 * gritas_b200_allreduce first.  The device accumulators are left untouched (Norm can be repeated).
 */
int gritas_b200_norm(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *stdv_2d,
                     double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *stdv_3d,
                     double *xcov_3d, double *nobs_3d);

/*
 * Replaces GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341; called by gritas.f:813-818 under -tch): the handle
 * accumulates three residual series (nr = 3).  bias (mean) and rtms as above, var_* receive the variances (no
 * square root, :1273 / :1320), xcov_*(...,1:3) the three-cornered-hat estimates -- combined exactly as the
 * reference's whole-array expressions do (:1287-1289, :1337-1339; missing values take part, the 2-D and the 3-D block
 * order slots 2 and 3 differently).  Returns GRITAS_B200_ERR_NR unless nr = 3.
 */
int gritas_b200_norm_3ch(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *var_2d,
                         double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *var_3d,
                         double *xcov_3d, double *nobs_3d);

/* Same finalize, results left in device memory (for callers that post-process on the GPU and for
 * timing the kernel without the D2H copy).  Pointers stay valid until the next norm/destroy. */
int gritas_b200_norm_device(gritas_b200_handle h, int nrmlz, int ntresh, const double **bias_2d,
                            const double **rtms_2d, const double **stdv_2d, const double **xcov_2d,
                            const double **nobs_2d, const double **bias_3d, const double **rtms_3d,
                            const double **stdv_3d, const double **xcov_3d, const double **nobs_3d);

/* The un-normalised sums in the reference's layout (what the host arrays hold between Accum and
 * Norm: gritas.f:812-818, -nonorm gritas.f:1508-1509).  xcov_* receive slot 1 only. */
int gritas_b200_get_raw(gritas_b200_handle h, double *bias_2d, double *rtms_2d, double *xcov_2d, double *nobs_2d,
                        double *bias_3d, double *rtms_3d, double *xcov_3d, double *nobs_3d);

/*
 * Scorecard of one synoptic time: replaces scores2_ + ave_nomiss (m_gritas_grids.f:1671-1826) and, together with
 * gritas_b200_reset + gritas_b200_accum on a scratch handle (nr = 1), the Accum + Norm of scores1_ (:1619-1634); the CSV
 * writing through the un-vendored m_obs_stats stays with the caller.  Finalizes with nrmlz = .true. on the device and
 * averages mean / rms / stdv over every region's boxes that are not AMISS, in the reference's summation order.
 *   regindx (2, nregs) first / last latitude row of each region, 1-based (init_, m_gritas_grids.f:163-175:
 *           gritas_b200_score_regions derives them from regbounds (2, nregs) exactly as init_ does);
 *   lon0, lon1  the longitude slot lonindx(:,1) = (1, im) (:180-181);
 *   collect_stats (nregs, km, l2d + l3d, 3) real(4), column-major: surface variables first, level slot 1 only
 *           (entries the reference leaves untouched are left untouched here).
 * Only that table is copied to the host; the handle's accumulation state is unchanged.
 */
int gritas_b200_scores(gritas_b200_handle h, int nregs, const int32_t *regindx, int lon0, int lon1, float *collect_stats);
int gritas_b200_score_regions(int jm, int nregs, const double *regbounds, int32_t *regindx);

/* Replaces GrITAS_MinMax_ (m_gritas_grids.f:1019-1167) together with gritas_b200_accum on a handle created with
 * GRITAS_B200_FLAG_MINMAX: minv (AMISS where empty), maxv (-AMISS where empty, gritas.f:997-1000 flips those)
 * and the counts, reference layout (im,jm,l2d) / (im,jm,km,l3d). */
int gritas_b200_get_minmax(gritas_b200_handle h, double *minv_2d, double *maxv_2d, double *nobs_2d, double *minv_3d,
                           double *maxv_3d, double *nobs_3d);

/* Replaces gritas_maskout (m_gritas_masks.F90:137-186, called at gritas.f:745): with a mask set, every accum
 * entry flags the observations whose box has maskfld(i,j) < 0.99 before accumulating (ob_flag = 1 on return).
 * maskfld is (im, jm) column-major on the handle's grid (ini_masks_ bins it to the target grid); NULL clears. */
int gritas_b200_set_mask(gritas_b200_handle h, const double *maskfld);

/* Adds caller-held un-normalised sums into the device accumulators (resuming from -nonorm output,
 * or honouring host grids that were not zero on the first Accum). */
int gritas_b200_add_raw(gritas_b200_handle h, const double *bias_2d, const double *rtms_2d, const double *xcov_2d,
                        const double *nobs_2d, const double *bias_3d, const double *rtms_3d, const double *xcov_3d,
                        const double *nobs_3d);

/*
 * Replaces the eight IndexSort calls of gritas.f:499-508 (GMAO_mpeu m_MergeSorts, the reference's largest host cost
 * per synoptic time): IndexSet + stable ascending sorts by qcexcl, lat, lon, lev, time, kt, ks, kx in this order, i.e.
 * the final order (kx, ks, kt, time, lev, lon, lat, qcexcl, original position).  `is` (nobs int32, host or device)
 * receives the permutation, 0- or 1-based (index_base; Fortran's `is` is 1-based); the column gathers of
 * gritas.f:510-523 stay with the caller.  The stable order is unique: the result is identical to the host sort for
 * NaN-free keys.  Needs a handle only for its device scratch and stream.
 */
int gritas_b200_presort(gritas_b200_handle h, int nobs, const int32_t *qcexcl, const double *lat, const double *lon,
                        const double *lev, const int32_t *time, const int32_t *kt, const int32_t *ks, const int32_t *kx,
                        int index_base, int32_t *is);

/*
 * The whole per-file preamble of gritas.f in one call: the pre-sort (gritas.f:499-508), the column gathers that follow it
 * (gritas.f:510-523), residual selection, QC filter and GrITAS_Accum_ as gritas_b200_accum_odsx.  `ods` holds the columns as
 * ODS_Get left them (unsorted), ks the sounding index column.  Nothing is sorted or gathered on the host.  In deterministic
 * mode the sums are formed in exactly the order the reference's sorted loop forms them.  ob_flag_out (may be NULL) is in
 * sorted order like the reference's ob_flag; is_out (may be NULL) receives the 1-based permutation `is`.
 */
int gritas_b200_accum_odsx_sorted(gritas_b200_handle h, int nobs, const gritas_b200_ods *ods, const int32_t *ks, int iopt,
                                  int options, int scale_res, int ioptn, int t1, int t2, double lona, double lonb,
                                  int x_passive, int x_pre_bad, int32_t *ob_flag_out, int32_t *is_out, double *bad_lev);

/* Multi-GPU (one process per GPU): partial count / sum / sum-of-squares grids are summed over
 * all ranks with one NCCL all-reduce before _norm.  The reference is single-process; this is new. */
int gritas_b200_nccl_unique_id(void *id128);  /* rank 0 fills 128 bytes; broadcast them */
int gritas_b200_comm_init(gritas_b200_handle h, int nranks, int rank, const void *id128);
int gritas_b200_allreduce(gritas_b200_handle h);
/* Multi-rank GrITAS_Norm in one collective call: the tile flush, an ncclReduce of the partial accumulators
 * to `root` and the finalize kernel are pipelined chunk by chunk (the reduction of chunk c overlaps the
 * accumulation of chunk c+1 and the finalize of chunk c-1).  Only the root's arrays are written (the
 * reference has one writer).  Same array meaning as gritas_b200_norm / _norm_device (outs10: the ten
 * pointers in the order bias_2d rtms_2d stdv_2d xcov_2d nobs_2d bias_3d rtms_3d stdv_3d xcov_3d nobs_3d). */
int gritas_b200_norm_reduce(gritas_b200_handle h, int root, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d,
                            double *stdv_2d, double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d,
                            double *stdv_3d, double *xcov_3d, double *nobs_3d);
int gritas_b200_norm_reduce_device(gritas_b200_handle h, int root, int nrmlz, int ntresh, const double **outs10);

/*
 * Sharded multi-rank GrITAS_Norm over peer memory (one node, up to 8 ranks; the reference is single-process, its
 * file loop gritas.f:389-393 runs on each rank's share f mod nranks and gritas.f:885-887 becomes this call).
 * Rank r finalizes 1/nranks of the rows (whole (j,l) rows, gritas_b200_slab) and its tile kernel pulls the parked
 * observations of those rows from every rank's tile lists through NVLink peer mappings -- the exchange step of the
 * path is fused into the accumulate + finalize kernel, no collective touches the sum grids.  Every rank writes only
 * its slab of the output arrays (same shapes and meaning as gritas_b200_norm): pass arrays in a shared-memory segment
 * (gritas_b200_host_register pins it) and the writer process sees the whole grids.
 *   peer_export  fills a GRITAS_B200_PEER_BLOB_BYTES blob (CUDA IPC handles of the lists / records + the settings that
 *                shape the exchange); gather the blobs of all ranks in rank order with whatever the host has
 *                (MPI_Allgather in the Fortran driver, torch.distributed in bench.py)
 *   peer_attach  maps the peers' buffers; fails on every rank alike if the ranks' settings differ
 *   norm_sharded collective over the ranks (call it on every rank); with a communicator (gritas_b200_comm_init) the two
 *                barriers it needs are stream-ordered one-element all-reduces, without one (several handles in one
 *                process) the caller orders the calls: gritas_b200_sync on every handle first
 * On the direct / deterministic path (no lists) the records of the slab are summed over the ranks in rank order
 * (reproducible) and finalized; that variant is one-shot (reset afterwards), like gritas_b200_allreduce.
 */
#define GRITAS_B200_PEER_BLOB_BYTES 512
int gritas_b200_peer_export(gritas_b200_handle h, void *blob);
int gritas_b200_peer_attach(gritas_b200_handle h, int nranks, int rank, const void *blobs);
int gritas_b200_slab(gritas_b200_handle h, int rank, int nranks, int *jl3_0, int *jl3_1, int *jl2_0, int *jl2_1);
int gritas_b200_norm_sharded(gritas_b200_handle h, int nrmlz, int ntresh, double *bias_2d, double *rtms_2d, double *stdv_2d,
                             double *xcov_2d, double *nobs_2d, double *bias_3d, double *rtms_3d, double *stdv_3d,
                             double *xcov_3d, double *nobs_3d);
int gritas_b200_norm_sharded_device(gritas_b200_handle h, int nrmlz, int ntresh, const double **outs10);
/* The sharded Norm exchanges parked ENTRIES (32 B each) while a rank parks fewer than ~3 per box, and sum PLANES
 * (8*(1+2nr+...) B per box: each rank exports its tiles' {n, sums}, the owner of a slab sums the ranks' planes through peer
 * memory) beyond -- every rank sees the same totals and decides alike (GRITAS_B200_PLANES_ABOVE overrides 3).  With a
 * communicator norm_sharded does the export itself; a host that drives several handles without one calls this on every
 * handle before the first norm_sharded. */
int gritas_b200_export_planes(gritas_b200_handle h);

/* Page-locks / releases caller-owned host memory (Fortran allocatables, a shared-memory segment): copies to and from
 * it are then direct DMA, no staging memcpy (replaces nothing in the reference; gritas.f:336-348 allocates the arrays). */
int gritas_b200_host_register(void *p, size_t bytes);
int gritas_b200_host_unregister(void *p);

/*
 * Channel cross-covariances ("a la Desroziers" with two residual series, three-cornered hat with three):
 *   gritas_b200_xcov_accum  replaces GrITAS_XCov_Accum_ (m_gritas_grids.f:667-848; called per synoptic time like
 *                           GrITAS_Accum).  The observations come as complete profiles of km consecutive entries; a
 *                           profile counts when each of the nchan active channels (achan = the achan(:, ichan_version)
 *                           column, m_gritas_grids.f:69-71; matched against NINT(lev)) is present, not flagged and in the
 *                           kx-kt table; such a profile adds the outer product of its residual vectors to xcov_lv.
 *                           ob_flag is inout as in the reference (set to 1 for active channels not in the table); the
 *                           reference's lat / lon arguments are unused there and absent here.  *nprof = accepted profiles.
 *                           rc GRITAS_B200_ERR_PROFILE: nobs is not a multiple of km, or a profile matched more than
 *                           nchan observations (both fatal in the reference).
 *   gritas_b200_xcov_norm   replaces GrITAS_XCov_Norm_ (m_gritas_grids.f:859-1008): means, covariances m/(m-1)(<xy> - ...)
 *                           (the symmetrised Desroziers form, or per series with tch) and, with tch and not self, the
 *                           three-cornered-hat combination.  Works on copies: the accumulators stay as they are.
 * Arrays, column-major: bias_lv (nchan, l3d, nr), xcov_lv (nchan, nchan, l3d, nr), nobs_lv (nchan, l3d); host or device.
 * The sums are formed profile by profile in input order with one multiply and one add per element: bit-identical to the
 * reference loop.  nr = 2 (tch = 0) or nr = 3 (tch = 1), as the reference demands (:737-745).
 */
typedef struct gritas_b200_xcov_ctx *gritas_b200_xcov_handle;
int gritas_b200_xcov_create(gritas_b200_xcov_handle *h, int km, int nchan, const int32_t *achan, int l3d, int nr, int tch,
                            const int32_t *kxkt_table, int kxmax, int ktmax, int device);
void gritas_b200_xcov_destroy(gritas_b200_xcov_handle h);
int gritas_b200_xcov_reset(gritas_b200_xcov_handle h);
int gritas_b200_xcov_accum(gritas_b200_xcov_handle h, int nobs, const double *lev, const int32_t *kx, const int32_t *kt,
                           const double *del, int32_t *ob_flag, int *nprof);
int gritas_b200_xcov_norm(gritas_b200_xcov_handle h, int nrmlz, int self, int ntresh, double *bias_lv, double *xcov_lv,
                          double *nobs_lv);
int gritas_b200_xcov_get_raw(gritas_b200_xcov_handle h, double *bias_lv, double *xcov_lv, double *nobs_lv);
/*
 * Level cross-covariances of conventional soundings: replaces GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059).  Profiles
 * are runs of equal sounding index ks; each gathers all observations carrying its index (an index that returns later is
 * processed once per run, as the reference does); levels from the GrITAS_Pint intervals p1 / p2; grid l = 1; the reference's
 * kx / kt / kxkt_table / ob_flag arguments are unused there and absent here.  Arrays as above with nchan = nlevs.  Sums are
 * formed in the reference's (n, m) order with separate multiplies and adds: bit-identical.  rc GRITAS_B200_ERR_LEVEL +
 * *bad_lev: 'could not find level' (:1990).  Norm / get_raw / reset / destroy: the gritas_b200_xcov_* entries above.
 */
int gritas_b200_xcov_prs_create(gritas_b200_xcov_handle *h, int nlevs, const double *p1, const double *p2, int l3d, int nr, int tch,
                                int device);
int gritas_b200_xcov_prs_accum(gritas_b200_xcov_handle h, int nobs, const double *lev, const int32_t *ks, const double *del, int *nprof,
                               double *bad_lev);
const char *gritas_b200_xcov_last_error(gritas_b200_xcov_handle h);
uint64_t gritas_b200_xcov_launch_count(gritas_b200_xcov_handle h);

/* Pinned host memory for the caller's column arrays (direct DMA, no staging copy). */
void *gritas_b200_host_alloc(size_t bytes);
void gritas_b200_host_free(void *p);

/* Introspection. */
const char *gritas_b200_last_error(gritas_b200_handle h);
void *gritas_b200_stream(gritas_b200_handle h);              /* cudaStream_t of the compute stream */
int gritas_b200_set_stream(gritas_b200_handle h, void *stream); /* run on a caller-owned stream */
int gritas_b200_tiled(gritas_b200_handle h);                 /* fast mode: 1 tiled accumulate, 0 direct, -1 undecided */
int gritas_b200_fused_norm(gritas_b200_handle h);            /* 1: the next Norm / get_raw consumes the tile lists itself (K2f: tile accumulate fused with finalize) */
uint64_t gritas_b200_launch_count(gritas_b200_handle h);     /* kernels launched so far */
uint64_t gritas_b200_h2d_bytes(gritas_b200_handle h);        /* bytes copied host->device so far */
uint64_t gritas_b200_d2h_bytes(gritas_b200_handle h);        /* bytes copied device->host so far */
int gritas_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GRITAS_B200_H */
