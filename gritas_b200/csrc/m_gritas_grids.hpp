// m_gritas_grids.hpp -- C++ host-side mirror of the reference's Fortran module m_gritas_grids for the
// gridding hot path, sitting directly on the C ABI (include/gritas_b200.h).
//
// The reference's host code is Fortran (src/Components/gritas/m_gritas_grids.f); this image has no Fortran
// compiler, so the caller-facing layer exists twice: fortran/m_gritas_b200.F90 (the real drop-in, source
// only) and this header, which keeps the same names, argument order, argument meaning and error behaviour
// so that a driver written against it reads like gritas.f.  Arrays are column-major and caller-owned.
//
//   grids_init     m_gritas_grids.f:105-190      GrITAS_Pint   :217-296
//   GrITAS_Accum   :307-478                      GrITAS_Norm   :490-656
//   GrITAS_KxKt    gritas_kxkt.f:10-119          gritas_perror gritas_etc.f:54-60 (print, exit(7))
#pragma once
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/gritas_b200.h"

namespace m_gritas_grids {

// module variables (m_gritas_grids.f:24-72)
inline int im = 0, jm = 0, km = 0, nlevs = 0;
inline std::vector<double> plevs;
inline double dLon = 0.0, dLat = 0.0;
constexpr double LonMin = -180.0, LatMin = -90.0, AMISS = 1.0e15;
inline gritas_b200_handle handle = nullptr;
inline int b200_mode = GRITAS_B200_MODE_FAST;

[[noreturn]] inline void gritas_perror(const char *msg) {   // gritas_etc.f:54-60
  std::printf("gritas: %s\n", msg);
  std::exit(7);
}

inline void grids_init(int im_, int jm_, int km_) {         // m_gritas_grids.f:105-190
  im = im_; jm = jm_; km = km_;
  dLon = 360.0 / im;                                        // :128
  dLat = 180.0 / (jm - 1);                                  // :129
}

inline void GrITAS_Pint(double *p1, double *p2) {           // m_gritas_grids.f:217-296
  static const char *msg[] = {"", "Pint: must have at least 1 vertical levels",
                              "Pint: pressure levels do not decrease monotonically", "Pint: cannot handle 0 hPa"};
  const int rc = gritas_b200_pint(nlevs, plevs.data(), p1, p2);
  if (rc) gritas_perror(msg[rc]);
}

inline void GrITAS_KxKt(int KTMAX, int KXMAX, int L2D_MAX, int L3D_MAX, int listsz, const int *kt_list, const int *kx_list,
                        const int *ktSurfAll, int nsurf, int *kxkt_table, int *list_2d, int &l2d, int *list_3d, int &l3d) {
  static const char *msg[] = {"", "KxKt: invalid kt - too large", "KxKt: invalid kt - negative",
                              "KxKt: invalid l2d - too large", "KxKt: invalid l3d - too large", "KxKt: invalid kx - too large"};
  const int rc = gritas_b200_kxkt(KTMAX, KXMAX, L2D_MAX, L3D_MAX, listsz, kt_list, kx_list, ktSurfAll, nsurf, kxkt_table,
                                  list_2d, list_3d, &l2d, &l3d);
  if (rc) gritas_perror(msg[rc]);
}

// Same argument list as GrITAS_Accum_ (m_gritas_grids.f:307-311).  The grids are accepted for signature parity
// and left untouched: the accumulators live on the GPU until GrITAS_Norm.  KXMAX / KTMAX are the m_odsmeta
// dimensions of kxkt_table.
inline void GrITAS_Accum(const double *lat, const double *lon, const double *lev, const int *kx, const int *kt,
                         const double *del, int nobs, int nr, const int *kxkt_table, int KXMAX, int KTMAX, const double *p1,
                         const double *p2, int *ob_flag, int l2d, double * /*bias_2d*/, double * /*rtms_2d*/,
                         double * /*xcov_2d*/, double * /*nobs_2d*/, int l3d, double * /*bias_3d*/, double * /*rtms_3d*/,
                         double * /*xcov_3d*/, double * /*nobs_3d*/) {
  if (nr > 3) gritas_perror("Accum: could not handle more then 2 residuals");   // :395-396
  if (!handle) {
    const int rc = gritas_b200_create(&handle, im, jm, km, nlevs, l2d, l3d, nr, p1, p2, kxkt_table, KXMAX, KTMAX, b200_mode, -1);
    if (rc) gritas_perror("Accum: cannot create the B200 context (no CUDA device?)");
  }
  double bad_lev = 0.0;
  const int rc = gritas_b200_accum(handle, nobs, lat, lon, lev, del, kx, kt, ob_flag, &bad_lev);
  if (rc == GRITAS_B200_ERR_LEVEL) {                                             // :458-459
    std::printf(" Accum: obs level is %g hPa\n", bad_lev);
    gritas_perror("Accum: could not find level");
  }
  if (rc) gritas_perror(gritas_b200_last_error(handle));
}

// Same argument list as GrITAS_Norm_ (m_gritas_grids.f:490-493); ntresh == nullptr: absent.
inline void GrITAS_Norm(int nr, bool nrmlz, int /*l2d*/, double *bias_2d, double *rtms_2d, double *stdv_2d, double *xcov_2d,
                        double *nobs_2d, int /*l3d*/, double *bias_3d, double *rtms_3d, double *stdv_3d, double *xcov_3d,
                        double *nobs_3d, const int *ntresh = nullptr) {
  if (nr > 2) gritas_perror("Accum: could not handle more then 2 residuals");   // :555 (sic)
  if (!handle) gritas_perror("Norm: called before Accum");
  const int rc = gritas_b200_norm(handle, nrmlz ? 1 : 0, ntresh ? *ntresh : -1, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,
                                  bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d);
  if (rc) gritas_perror(gritas_b200_last_error(handle));
}

inline void GrITAS_Clean() {                                                      // gritas.f:2017-2026
  if (handle) gritas_b200_destroy(handle);
  handle = nullptr;
}

}  // namespace m_gritas_grids
