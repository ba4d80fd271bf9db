"""ctypes front-end of the CPU oracle (oracle/gritas_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the gritas_b200 package.
Pinned against the reference's own routines translated to C (oracle/ref.py, tests/test_oracle_vs_ref.py) for the
gridding path; the rest is unpinned (see the header of gritas_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libgritas_oracle.so")

AMISS = 1.0e15  # m_gritas_grids.f:72

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="F_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="F_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="F_CONTIGUOUS")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gritas_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_pint.argtypes = [C.c_int, _f64p, _f64p, _f64p]
        L.orc_pint.restype = C.c_int
        L.orc_kxkt.argtypes = [C.c_int] * 5 + [_i32p, _i32p, _i32p, C.c_int, _i32p, _i32p, _i32p,
                                               C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_kxkt.restype = C.c_int
        L.orc_presort.argtypes = [C.c_int, _i32p, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.orc_presort.restype = C.c_int
        L.orc_select_del.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.orc_select_del.restype = C.c_int
        L.orc_select_del_all.argtypes = [C.c_int] * 5 + [_f64p] * 6
        L.orc_select_del_all.restype = C.c_int
        L.orc_filter.argtypes = [C.c_int, _i32p, _i32p, _f64p, C.c_int, C.c_int, C.c_int, C.c_double,
                                 C.c_double, C.c_int, C.c_int, _i32p]
        L.orc_filter.restype = None
        L.orc_accum.argtypes = ([C.c_int] * 6 + [_f64p, _f64p, _f64p, _i32p, _i32p, _f64p, _i32p, C.c_int,
                                                 C.c_int, _f64p, _f64p, _i32p,
                                                 C.c_int, _f64p, _f64p, _f64p, _f64p,
                                                 C.c_int, _f64p, _f64p, _f64p, _f64p,
                                                 C.POINTER(C.c_double), C.POINTER(C.c_int)])
        L.orc_accum.restype = C.c_int
        L.orc_box_keys.argtypes = ([C.c_int] * 5 + [_f64p, _f64p, _f64p, _i32p, _i32p, _i32p, C.c_int, C.c_int,
                                                    _f64p, _f64p, _i32p, C.c_int, C.c_int, _i64p])
        L.orc_box_keys.restype = None
        L.orc_minmax.argtypes = ([C.c_int] * 6 + [_f64p, _f64p, _f64p, _i32p, _i32p, _f64p, _i32p, C.c_int, C.c_int, _f64p, _f64p,
                                                  _i32p, C.c_int, _f64p, _f64p, _f64p, C.c_int, _f64p, _f64p, _f64p,
                                                  C.POINTER(C.c_double), C.POINTER(C.c_int)])
        L.orc_minmax.restype = C.c_int
        L.orc_maskout.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _f64p, _i32p]
        L.orc_maskout.restype = C.c_int
        L.orc_norm.argtypes = ([C.c_int] * 6 + [C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p,
                                                C.c_int, _f64p, _f64p, _f64p, _f64p, _f64p, C.c_int])
        L.orc_norm.restype = C.c_int
        L.orc_norm_3ch.argtypes = L.orc_norm.argtypes
        L.orc_norm_3ch.restype = C.c_int
        _lib = L
    return _lib


def grid_init(im: int, jm: int):
    """m_gritas_grids.f:128-129."""
    return 360.0 / im, 180.0 / (jm - 1)


def pint(plevs):
    """GrITAS_Pint_ (m_gritas_grids.f:217-296) -> (p1, p2); raises on the gritas_perror cases."""
    plevs = np.ascontiguousarray(plevs, dtype=np.float64)
    n = plevs.size
    p1 = np.zeros(max(n, 1))
    p2 = np.zeros(max(n, 1))
    rc = lib().orc_pint(n, plevs if n else np.zeros(1), p1, p2)
    if rc:
        raise ValueError({1: "Pint: must have at least 1 vertical levels",
                          2: "Pint: pressure levels do not decrease monotonically",
                          3: "Pint: cannot handle 0 hPa"}[rc])
    return p1[:n], p2[:n]


def kxkt(ktmax, kxmax, kt_list, kx_lists, kt_surf=(1, 2, 3), l2d_max=20, l3d_max=22):
    """GrITAS_KxKt (gritas_kxkt.f:10-119).  kx_lists: one sequence of kx per list row.
    Returns (table[kxmax,ktmax] F-order int32, list_2d, list_3d, l2d, l3d)."""
    listsz = len(kt_list)
    kxl = np.full((kxmax, max(listsz, 1)), -1, dtype=np.int32, order="F")
    for j, ks in enumerate(kx_lists):
        kxl[: len(ks), j] = np.asarray(ks, dtype=np.int32)
    table = np.zeros((kxmax, ktmax), dtype=np.int32, order="F")
    l2 = np.zeros(max(l2d_max, 1), dtype=np.int32)
    l3 = np.zeros(max(l3d_max, 1), dtype=np.int32)
    a, b = C.c_int(0), C.c_int(0)
    ktl = np.asarray(kt_list, dtype=np.int32) if listsz else np.zeros(1, np.int32)
    surf = np.asarray(kt_surf, dtype=np.int32) if len(kt_surf) else np.zeros(1, np.int32)
    rc = lib().orc_kxkt(ktmax, kxmax, l2d_max, l3d_max, listsz, ktl, kxl, surf, len(kt_surf), table, l2, l3,
                        C.byref(a), C.byref(b))
    if rc:
        raise ValueError({1: "KxKt: invalid kt - too large", 2: "KxKt: invalid kt - negative",
                          3: "KxKt: invalid l2d - too large", 4: "KxKt: invalid l3d - too large",
                          5: "KxKt: invalid kx - too large"}[rc])
    return table, l2[: a.value].copy(), l3[: b.value].copy(), a.value, b.value


def presort(qcexcl, lat, lon, lev, time, kt, ks, kx):
    """gritas.f:499-508 -> 0-based permutation."""
    n = len(lat)
    out = np.zeros(max(n, 1), dtype=np.int32)
    rc = lib().orc_presort(n, *_c(np.int32, qcexcl), *_c(np.float64, lat, lon, lev),
                           *_c(np.int32, time, kt, ks, kx), out)
    if rc:
        raise MemoryError
    return out[:n]


def _c(dt, *arrs):
    out = []
    for a in arrs:
        a = np.ascontiguousarray(a, dtype=dt)
        if a.size == 0:
            a = np.zeros(1, dtype=dt)
        out.append(a)
    return out


def select_del(iopt, omf, oma, obs=None, xvec=None, scale_res=0):
    """gritas.f:562-587, column 1."""
    n = len(omf)
    obs = np.zeros(n) if obs is None else obs
    xvec = np.ones(n) if xvec is None else xvec
    d = np.zeros(max(n, 1))
    rc = lib().orc_select_del(n, iopt, scale_res, *_c(np.float64, omf, oma, obs, xvec), d)
    if rc:
        raise ValueError("iopt not covered by the oracle")
    return d[:n]


def select_del_all(iopt, omf, oma, obs, xvec, xm, tch=False, self_=False, scale_res=0):
    """gritas.f:562-704 (+ :738) for every single-ODS residual option -> del (nobs, nr) F-order."""
    n = len(omf)
    d = np.zeros((max(n, 1), 3), order="F")
    nr = lib().orc_select_del_all(n, iopt, int(tch), int(self_), scale_res, *_c(np.float64, omf, oma, obs, xvec, xm),
                                  d.reshape(-1, order="F"))
    if nr == 0:
        raise ValueError("iopt not covered by the oracle")
    return np.asfortranarray(d[:n, :nr])


def filter_flags(qcexcl, time, lon, ioptn=1, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7, x_pre_bad=1):
    """gritas.f:710-739 -> ob_flag."""
    n = len(qcexcl)
    flag = np.zeros(max(n, 1), dtype=np.int32)
    lib().orc_filter(n, *_c(np.int32, qcexcl, time), *_c(np.float64, lon), ioptn, t1, t2, lona, lonb,
                     x_passive, x_pre_bad, flag)
    return flag[:n]


class LevelMiss(Exception):
    """gritas_perror('Accum: could not find level') (m_gritas_grids.f:458-459)."""

    def __init__(self, lev, n):
        super().__init__(f"Accum: obs level is {lev} hPa; Accum: could not find level (obs {n})")
        self.lev = lev
        self.n = n


class Grids:
    """The reference's allocatable grids (gritas.f:336-363), zero-initialised, column-major."""

    def __init__(self, im, jm, km, l2d, l3d, nr):
        self.im, self.jm, self.km, self.l2d, self.l3d, self.nr = im, jm, km, l2d, l3d, nr
        z = lambda *s: np.zeros(s, dtype=np.float64, order="F")
        a2, a3 = max(l2d, 1), max(l3d, 1)
        self.bias_2d, self.rtms_2d, self.stdv_2d, self.xcov_2d = (z(im, jm, a2, nr) for _ in range(4))
        self.nobs_2d = z(im, jm, a2)
        self.bias_3d, self.rtms_3d, self.stdv_3d, self.xcov_3d = (z(im, jm, km, a3, nr) for _ in range(4))
        self.nobs_3d = z(im, jm, km, a3)

    def names(self):
        return ["bias_2d", "rtms_2d", "stdv_2d", "xcov_2d", "nobs_2d",
                "bias_3d", "rtms_3d", "stdv_3d", "xcov_3d", "nobs_3d"]


def accum(g: Grids, nlevs, lat, lon, lev, kx, kt, del_, kxkt_table, p1, p2, ob_flag):
    """GrITAS_Accum_ (m_gritas_grids.f:307-478).  del_ is (nobs, nr) F-order; ob_flag int32 inout."""
    nobs = len(lat)
    if nobs == 0:
        return
    del_ = np.asfortranarray(np.asarray(del_, dtype=np.float64).reshape(nobs, -1, order="F"))
    nr = del_.shape[1]
    assert nr == g.nr
    assert ob_flag.dtype == np.int32
    kxmax, ktmax = kxkt_table.shape
    bad, badn = C.c_double(0.0), C.c_int(-1)
    if nobs == 0:
        return
    rc = lib().orc_accum(g.im, g.jm, g.km, nlevs, nobs, nr, *_c(np.float64, lat, lon, lev),
                         *_c(np.int32, kx, kt), del_, np.asfortranarray(kxkt_table, dtype=np.int32), kxmax, ktmax,
                         *_c(np.float64, p1, p2), ob_flag,
                         g.l2d, g.bias_2d, g.rtms_2d, g.xcov_2d, g.nobs_2d,
                         g.l3d, g.bias_3d, g.rtms_3d, g.xcov_3d, g.nobs_3d, C.byref(bad), C.byref(badn))
    if rc == 7:
        raise LevelMiss(bad.value, badn.value)
    if rc:
        raise ValueError("Accum: could not handle more then 2 residuals")


def box_keys(im, jm, km, nlevs, lat, lon, lev, kx, kt, kxkt_table, p1, p2, ob_flag, l2d, l3d):
    nobs = len(lat)
    key = np.zeros(max(nobs, 1), dtype=np.int64)
    kxmax, ktmax = kxkt_table.shape
    lib().orc_box_keys(im, jm, km, nlevs, nobs, *_c(np.float64, lat, lon, lev), *_c(np.int32, kx, kt),
                       np.asfortranarray(kxkt_table, dtype=np.int32), kxmax, ktmax, *_c(np.float64, p1, p2),
                       *_c(np.int32, ob_flag), l2d, l3d, key)
    return key[:nobs]


def norm(g: Grids, nlevs, nrmlz=True, ntresh=0):
    """GrITAS_Norm_ (m_gritas_grids.f:490-656), in place."""
    rc = lib().orc_norm(g.im, g.jm, g.km, nlevs, g.nr, int(bool(nrmlz)),
                        g.l2d, g.bias_2d, g.rtms_2d, g.stdv_2d, g.xcov_2d, g.nobs_2d,
                        g.l3d, g.bias_3d, g.rtms_3d, g.stdv_3d, g.xcov_3d, g.nobs_3d, ntresh)
    if rc:
        raise ValueError("Norm: could not handle more then 2 residuals")


def norm_3ch(g: Grids, nlevs, nrmlz=True, ntresh=0):
    """GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341), in place: g.stdv_* receive the variances, g.xcov_* the 3CH estimates."""
    rc = lib().orc_norm_3ch(g.im, g.jm, g.km, nlevs, g.nr, int(bool(nrmlz)),
                            g.l2d, g.bias_2d, g.rtms_2d, g.stdv_2d, g.xcov_2d, g.nobs_2d,
                            g.l3d, g.bias_3d, g.rtms_3d, g.stdv_3d, g.xcov_3d, g.nobs_3d, ntresh)
    if rc:
        raise ValueError("GrITAS_3CH_Norm_ could not handle more then 2 residuals")


def minmax(g: Grids, nlevs, lat, lon, lev, kx, kt, del_, kxkt_table, p1, p2, ob_flag):
    """GrITAS_MinMax_ (m_gritas_grids.f:1019-1167): min in g.bias_*, max in g.stdv_*, counts in g.nobs_* -- the arrays
    gritas.f:770-775 passes; start them at AMISS / -AMISS (gritas.f:364-366) with minmax_init."""
    nobs = len(lat)
    if nobs == 0:
        return
    d = np.asfortranarray(np.asarray(del_, dtype=np.float64).reshape(nobs, -1, order="F"))
    kxmax, ktmax = kxkt_table.shape
    bad, badn = C.c_double(0.0), C.c_int(-1)
    rc = lib().orc_minmax(g.im, g.jm, g.km, nlevs, nobs, d.shape[1], *_c(np.float64, lat, lon, lev), *_c(np.int32, kx, kt),
                          d, np.asfortranarray(kxkt_table, dtype=np.int32), kxmax, ktmax, *_c(np.float64, p1, p2), ob_flag,
                          g.l2d, g.bias_2d, g.stdv_2d, g.nobs_2d, g.l3d, g.bias_3d, g.stdv_3d, g.nobs_3d,
                          C.byref(bad), C.byref(badn))
    if rc == 7:
        raise LevelMiss(bad.value, badn.value)
    if rc:
        raise ValueError("MinMax: could not handle more then 1 field")


def minmax_init(g: Grids):
    """gritas.f:364-366."""
    for a in (g.bias_2d, g.bias_3d):
        a[...] = AMISS
    for a in (g.stdv_2d, g.stdv_3d):
        a[...] = -AMISS


def maskout(im, jm, lat, lon, maskfld, qcx):
    """m_gritas_masks.F90:137-186; qcx int32 inout, returns the number of observations excluded."""
    return lib().orc_maskout(im, jm, len(lat), *_c(np.float64, lat, lon), np.asfortranarray(maskfld, dtype=np.float64), qcx)


# ---- channel / level cross-covariances (m_gritas_grids.f:667-1008, 1862-2059) ------------------------------
_XC_ERR = {1: "improper dim settings", 2: "TCH setting inconsistent w/ dims", 3: "Desroziers setting inconsistent w/ dims",
           4: "XCov_Accum: not all profiles complete, aborting ...", 5: "XCov-Accum: error in channel matching"}


def _xc_protos():
    L = lib()
    if getattr(L, "_xc_ready", False):
        return L
    L.orc_xcov_accum.argtypes = [C.c_int, C.c_int, _i32p, C.c_int, C.c_int, _f64p, _i32p, _i32p, _f64p, _i32p, C.c_int, C.c_int,
                                 _i32p, C.c_int, _f64p, _f64p, _f64p, C.POINTER(C.c_int), C.c_int]
    L.orc_xcov_accum.restype = C.c_int
    L.orc_xcov_norm.argtypes = [C.c_int] * 6 + [_f64p, _f64p, _f64p, C.c_int]
    L.orc_xcov_norm.restype = C.c_int
    L.orc_xcov_prs_accum.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _i32p, _f64p, _f64p, _f64p, C.c_int, _f64p, _f64p, _f64p,
                                     C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_double)]
    L.orc_xcov_prs_accum.restype = C.c_int
    L._xc_ready = True
    return L


def xcov_alloc(ndim, l3d, nr):
    return (np.zeros((ndim, l3d, nr), order="F"), np.zeros((ndim, ndim, l3d, nr), order="F"), np.zeros((ndim, l3d), order="F"))


def xcov_accum(km, achan, lev, kx, kt, del_, kxkt_table, ob_flag, l3d, bias_lv, xcov_lv, nobs_lv, tch):
    """GrITAS_XCov_Accum_ (m_gritas_grids.f:667-848).  ob_flag int32 inout.  Returns nprof."""
    L = _xc_protos()
    table = np.asfortranarray(kxkt_table, dtype=np.int32)
    d = np.asfortranarray(del_, dtype=np.float64)
    nprof = C.c_int(0)
    rc = L.orc_xcov_accum(km, len(achan), np.ascontiguousarray(achan, dtype=np.int32), len(lev), d.shape[1],
                          *_c(np.float64, lev), *_c(np.int32, kx, kt), d, table, table.shape[0], table.shape[1], ob_flag, l3d,
                          bias_lv, xcov_lv, nobs_lv, C.byref(nprof), 1 if tch else 0)
    if rc:
        raise ValueError(_XC_ERR[rc])
    return nprof.value


def xcov_norm(nr, nrmlz, tch, self_, l3d, bias_lv, xcov_lv, nobs_lv, ntresh=0):
    """GrITAS_XCov_Norm_ (m_gritas_grids.f:859-1008), in place."""
    rc = _xc_protos().orc_xcov_norm(bias_lv.shape[0], nr, 1 if nrmlz else 0, 1 if tch else 0, 1 if self_ else 0, l3d, bias_lv,
                                    xcov_lv, nobs_lv, ntresh)
    if rc:
        raise ValueError(_XC_ERR[rc])


def xcov_prs_accum(nlevs, lev, ks, del_, p1, p2, l3d, bias_lv, xcov_lv, nobs_lv, tch):
    """GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059).  Returns nprof; raises LevelMiss."""
    L = _xc_protos()
    d = np.asfortranarray(del_, dtype=np.float64)
    nprof, bad = C.c_int(0), C.c_double(0.0)
    rc = L.orc_xcov_prs_accum(nlevs, len(lev), d.shape[1], *_c(np.float64, lev), *_c(np.int32, ks), d, *_c(np.float64, p1, p2), l3d,
                              bias_lv, xcov_lv, nobs_lv, C.byref(nprof), 1 if tch else 0, C.byref(bad))
    if rc == 7:
        raise LevelMiss(bad.value, -1)
    if rc:
        raise ValueError(_XC_ERR[rc])
    return nprof.value


# ---- scorecard (m_gritas_grids.f:163-175, 1671-1826) ------------------------------------------------------------
REGBOUNDS = ((-90.0, 90.0), (20.0, 80.0), (-20.0, 20.0), (-80.0, -20.0))   # global, n.hem, tropics, s.hem (:58, :146-157)


def score_regions(jm, regbounds=REGBOUNDS):
    L = lib()
    L.orc_score_regions.argtypes = [C.c_int, C.c_int, _f64p, _i32p]
    L.orc_score_regions.restype = None
    rb = np.asfortranarray(np.array(regbounds, dtype=np.float64).T)
    out = np.zeros((2, rb.shape[1]), np.int32, order="F")
    L.orc_score_regions(jm, rb.shape[1], rb, out)
    return out


def scores2(g, regindx):
    """scores2_ on finalized grids (nr = 1) -> collect_stats(mregs, km, l2d + l3d, 3) float32."""
    L = lib()
    f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="F_CONTIGUOUS")
    L.orc_scores2.argtypes = [C.c_int] * 6 + [_i32p, _i32p] + [_f64p] * 8 + [f32p]
    L.orc_scores2.restype = None
    nreg = regindx.shape[1]
    out = np.zeros((nreg, g.km, g.l2d + g.l3d, 3), np.float32, order="F")
    lon = np.array([1, g.im], np.int32)
    sl = lambda a: np.asfortranarray(a[..., 0]) if a.ndim in (4, 5) else a
    L.orc_scores2(g.im, g.jm, g.km, g.l2d, g.l3d, nreg, np.asfortranarray(regindx, dtype=np.int32), lon,
                  sl(g.bias_2d), sl(g.rtms_2d), sl(g.stdv_2d), g.nobs_2d, sl(g.bias_3d), sl(g.rtms_3d), sl(g.stdv_3d), g.nobs_3d, out)
    return out
