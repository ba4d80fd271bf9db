"""ctypes front-end of oracle/_ref/libgritas_ref.so: the reference's OWN Fortran routines (m_gritas_grids.f: GrITAS_Pint_,
GrITAS_Accum_, GrITAS_Norm_, GrITAS_MinMax_, GrITAS_3CH_Norm_ and the two statements of init_ that define the mesh),
translated to C statement by statement by oracle/f2c_subset.py and compiled with gcc.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_golden.py).  The library is built in THIS container from the
sources where they lie under /root/reference (`build()`; oracle/_ref/ is git-ignored and travels to the GPU box as a
built artefact); where neither the sources nor a built library exist, `available()` is False and the tests that need
it are skipped.  Same call signatures as oracle/oracle.py, same Grids class.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from . import oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_LIB = os.path.join(_DIR, "libgritas_ref.so")
_GEN = os.path.join(_DIR, "gritas_ref_gen.c")
REF_ROOT = "/root/reference/src"
REF_FILES = [os.path.join(REF_ROOT, f) for f in ("Components/gritas/m_gritas_grids.f", "Components/gritas/gritas_kxkt.f",
                                                 "Components/gritas/m_gritas_masks.F90", "Applications/GrITAS_App/gritas.f")]
CFLAGS = ["-O2", "-march=x86-64-v2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-Wall", "-Wno-unused-label",
          "-Wno-unused-variable", "-Wno-unused-but-set-variable"]

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="F_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="F_CONTIGUOUS")


def build(force: bool = False) -> str | None:
    """translate + compile when the reference sources are here; otherwise use what is already built (or nothing)"""
    if all(os.path.exists(f) for f in REF_FILES):
        deps = REF_FILES + [os.path.join(_HERE, "f2c_subset.py"), os.path.join(_HERE, "ref_harness.c")]
        if force or not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps):
            os.makedirs(_DIR, exist_ok=True)
            subprocess.check_call([sys.executable, os.path.join(_HERE, "f2c_subset.py"), REF_ROOT, _GEN])
            subprocess.check_call(["gcc"] + CFLAGS + ["-shared", "-o", _LIB, os.path.join(_HERE, "ref_harness.c"), "-lm"], cwd=_HERE)
            if not os.environ.get("GRITAS_KEEP_REF_C"):
                os.remove(_GEN)        # the translated text is an intermediate: only the compiled library stays (and travels)
    return _LIB if os.path.exists(_LIB) else None


def available() -> bool:
    try:
        return build() is not None
    except (subprocess.CalledProcessError, OSError):
        return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/libgritas_ref.so is not built and /root/reference is not here")
        L = C.CDLL(_LIB)
        L.ref_init.argtypes = [C.c_int] * 4 + [_f64p, C.c_int, C.c_int]
        L.ref_pint.argtypes = [_f64p, _f64p]
        grids2, grids3 = [_f64p] * 4, [_f64p] * 4
        L.ref_accum.argtypes = ([_f64p, _f64p, _f64p, _i32p, _i32p, _f64p, C.c_int, C.c_int, _i32p, _f64p, _f64p, _i32p, C.c_int]
                                + grids2 + [C.c_int] + grids3)
        L.ref_minmax.argtypes = ([_f64p, _f64p, _f64p, _i32p, _i32p, _f64p, C.c_int, C.c_int, _i32p, _f64p, _f64p, _i32p, C.c_int]
                                 + [_f64p] * 3 + [C.c_int] + [_f64p] * 3)
        for f in (L.ref_norm, L.ref_norm_3ch):
            f.argtypes = [C.c_int, C.c_int, C.c_int] + [_f64p] * 5 + [C.c_int] + [_f64p] * 5 + [C.c_int, C.c_int]
        L.ref_dlon.restype = L.ref_dlat.restype = C.c_double
        L.ref_set_surface_kts.argtypes = [_i32p, C.c_int]
        L.ref_kxkt.argtypes = [C.c_int] * 5 + [_i32p] * 5 + [C.POINTER(C.c_int)] * 2
        L.ref_select.argtypes = [C.c_int] * 10 + [C.POINTER(C.c_int), _f64p] + [_f64p] * 5 + [_i32p] + [_f64p] * 3
        L.ref_filter.argtypes = [C.c_int] * 4 + [C.c_double] * 2 + [C.c_int] * 2 + [_i32p, _i32p, _f64p, _i32p]
        L.ref_jo_scale.argtypes = [C.c_int] * 3 + [_f64p]
        L.ref_set_channels.argtypes = [_i32p, C.c_int, C.c_int]
        L.ref_ave_nomiss.argtypes = [_f64p, _f64p, C.c_int, C.c_int, _i32p, _i32p]
        L.ref_ave_nomiss.restype = C.c_double
        L.ref_region_rows.argtypes = [C.c_int, _f64p, _i32p]
        L.ref_scores2.argtypes = [np.ctypeslib.ndpointer(dtype=np.float32, flags="F_CONTIGUOUS")] + [C.c_int] * 7 + [_i32p] + [_f64p] * 8
        L.ref_maskout.argtypes = [C.c_int, C.c_int, C.c_int, _f64p, _f64p, _f64p, _i32p]
        L.ref_presort.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p, _i32p, _i32p, _i32p, _i32p]
        L.ref_xcov_accum.argtypes = ([_f64p] * 3 + [_i32p, _i32p, _f64p, C.c_int, C.c_int, _i32p, _i32p, C.c_int, _f64p] + [C.c_int] * 3
                                     + [_f64p, _f64p, C.POINTER(C.c_int), C.c_int])
        L.ref_xcov_norm.argtypes = ([C.c_int] * 5 + [_f64p] + [C.c_int] * 3 + [_f64p] + [C.c_int] * 4 + [_f64p] + [C.c_int] * 2
                                    + [C.c_int, C.c_int])
        L.ref_xcov_prs_accum.argtypes = ([_f64p, _i32p, _i32p, _i32p, _f64p, C.c_int, C.c_int, _i32p, _f64p, _f64p, _i32p, C.c_int, _f64p]
                                         + [C.c_int] * 3 + [_f64p] + [C.c_int] * 4 + [_f64p] + [C.c_int] * 2 + [C.POINTER(C.c_int), C.c_int])
        _lib = L
    return _lib


class RefError(Exception):
    """gritas_perror was called (the reference prints and exits with status 7); `line` is the source line of the call"""

    def __init__(self, line):
        super().__init__(f"gritas_perror called at m_gritas_grids.f:{line}")
        self.line = line


def _check(rc):
    if rc:
        raise RefError(lib().ref_error_line())


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def init(im, jm, km, nlevs, plevs, kxmax, ktmax):
    """module state as grids_init + the rc file leave it (m_gritas_grids.f:105-190)"""
    pl = np.zeros(max(km, nlevs, 1))
    pl[:nlevs] = np.asarray(plevs, dtype=np.float64)[:nlevs]
    lib().ref_init(im, jm, km, nlevs, pl, kxmax, ktmax)


def mesh():
    return lib().ref_dlon(), lib().ref_dlat()


def pint(nlevs):
    p1, p2 = np.zeros(nlevs), np.zeros(nlevs)
    _check(lib().ref_pint(p1, p2))
    return p1, p2


def accum(g: oracle.Grids, lat, lon, lev, kx, kt, del_, kxkt_table, p1, p2, ob_flag):
    nobs = len(lat)
    d = np.asfortranarray(np.asarray(del_, dtype=np.float64).reshape(nobs, -1, order="F"))
    assert ob_flag.dtype == np.int32 and d.shape[1] == g.nr
    _check(lib().ref_accum(_f(lat), _f(lon), _f(lev), _i(kx), _i(kt), d, nobs, g.nr, np.asfortranarray(kxkt_table, dtype=np.int32),
                           _f(p1), _f(p2), ob_flag, g.l2d, g.bias_2d, g.rtms_2d, g.xcov_2d, g.nobs_2d,
                           g.l3d, g.bias_3d, g.rtms_3d, g.xcov_3d, g.nobs_3d))


def minmax(g: oracle.Grids, lat, lon, lev, kx, kt, del_, kxkt_table, p1, p2, ob_flag):
    nobs = len(lat)
    d = np.asfortranarray(np.asarray(del_, dtype=np.float64).reshape(nobs, -1, order="F"))
    _check(lib().ref_minmax(_f(lat), _f(lon), _f(lev), _i(kx), _i(kt), d, nobs, d.shape[1], np.asfortranarray(kxkt_table, dtype=np.int32),
                            _f(p1), _f(p2), ob_flag, g.l2d, g.bias_2d, g.stdv_2d, g.nobs_2d, g.l3d, g.bias_3d, g.stdv_3d, g.nobs_3d))


def norm(g: oracle.Grids, nrmlz=True, ntresh=None):
    """ntresh=None: the optional argument is absent, as in gritas.f:885-887"""
    _check(lib().ref_norm(g.nr, int(bool(nrmlz)), g.l2d, g.bias_2d, g.rtms_2d, g.stdv_2d, g.xcov_2d, g.nobs_2d,
                          g.l3d, g.bias_3d, g.rtms_3d, g.stdv_3d, g.xcov_3d, g.nobs_3d, ntresh or 0, int(ntresh is not None)))


def norm_3ch(g: oracle.Grids, nrmlz=True, ntresh=None):
    _check(lib().ref_norm_3ch(g.nr, int(bool(nrmlz)), g.l2d, g.bias_2d, g.rtms_2d, g.stdv_2d, g.xcov_2d, g.nobs_2d,
                              g.l3d, g.bias_3d, g.rtms_3d, g.stdv_3d, g.xcov_3d, g.nobs_3d, ntresh or 0, int(ntresh is not None)))


def kxkt(ktmax, kxmax, kt_list, kx_lists, kt_surf=(1, 2, 3), l2d_max=20, l3d_max=22):
    """GrITAS_KxKt (gritas_kxkt.f:10-119), same return value as oracle.kxkt"""
    listsz = len(kt_list)
    kxl = np.full((kxmax, max(listsz, 1)), -1, dtype=np.int32, order="F")
    for j, ks in enumerate(kx_lists):
        kxl[: len(ks), j] = np.asarray(ks, dtype=np.int32)
    table = np.zeros((kxmax, ktmax), dtype=np.int32, order="F")
    l2, l3 = np.zeros(max(l2d_max, 1), np.int32), np.zeros(max(l3d_max, 1), np.int32)
    a, b = C.c_int(0), C.c_int(0)
    lib().ref_set_surface_kts(np.asarray(list(kt_surf) or [0], dtype=np.int32), len(kt_surf))
    _check(lib().ref_kxkt(ktmax, kxmax, l2d_max, l3d_max, listsz, np.asarray(list(kt_list) or [0], dtype=np.int32), kxl, table,
                          l2, l3, C.byref(a), C.byref(b)))
    return table, l2[: a.value].copy(), l3[: b.value].copy(), a.value, b.value


def select(iopt, nr, omf, oma, obs, xvec, xm, qcexcl, tch=False, self_=False, scale_res=0, xcov=False, pxcov=False,
           meanres=False, m_obs=None, m_omf=None, m_oma=None, x_passive=7):
    """gritas.f:562-704: the residual selection chain.  Returns (del (nobs, nr) F-order, qcexcl after it, nstat or None)."""
    n = len(omf)
    d = np.zeros((n, nr), order="F")
    qc = np.ascontiguousarray(qcexcl, dtype=np.int32).copy()
    z = lambda a: _f(a if a is not None else np.zeros(n))
    nstat = C.c_int(-1)
    _check(lib().ref_select(iopt, scale_res, int(xcov), int(pxcov), int(tch), int(self_), int(meanres), n, nr, x_passive,
                            C.byref(nstat), d, _f(omf), _f(oma), _f(obs), _f(xvec), _f(xm), qc, z(m_obs), z(m_omf), z(m_oma)))
    return d, qc, (nstat.value if nstat.value >= 0 else None)


def jo_scale(iopt, d):
    """gritas.f:738"""
    d = np.asfortranarray(d)
    _check(lib().ref_jo_scale(iopt, d.shape[0], d.shape[1], d))
    return d


def filter_flags(qcexcl, time, lon, ioptn=1, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7, x_pre_bad=1):
    """gritas.f:712-735 -> ob_flag"""
    n = len(qcexcl)
    flag = np.full(n, -9, dtype=np.int32)
    _check(lib().ref_filter(n, ioptn, t1, t2, lona, lonb, x_passive, x_pre_bad, _i(qcexcl), _i(time), _f(lon), flag))
    return flag


def xcov_accum(achan, lev, kx, kt, del_, kxkt_table, ob_flag, l3d, bias_lv, xcov_lv, nobs_lv, tch):
    """GrITAS_XCov_Accum_ (m_gritas_grids.f:667-848); the profile length is the module's km (ref.init).  Returns nprof."""
    d = np.asfortranarray(del_, dtype=np.float64)
    n = len(lev)
    lib().ref_set_channels(_i(achan), len(achan), 1)
    nprof = C.c_int(0)
    z = np.zeros(n)
    _check(lib().ref_xcov_accum(z, z, _f(lev), _i(kx), _i(kt), d, n, d.shape[1], np.asfortranarray(kxkt_table, dtype=np.int32), ob_flag,
                                l3d, bias_lv, *bias_lv.shape, xcov_lv, nobs_lv, C.byref(nprof), int(bool(tch))))
    return nprof.value


def xcov_norm(nr, nrmlz, tch, self_, l3d, bias_lv, xcov_lv, nobs_lv, ntresh=None):
    """GrITAS_XCov_Norm_ (m_gritas_grids.f:859-1008), in place"""
    _check(lib().ref_xcov_norm(nr, int(bool(nrmlz)), int(bool(tch)), int(bool(self_)), l3d, bias_lv, *bias_lv.shape, xcov_lv,
                               *xcov_lv.shape, nobs_lv, *nobs_lv.shape, ntresh or 0, int(ntresh is not None)))


def xcov_prs_accum(lev, ks, del_, p1, p2, l3d, bias_lv, xcov_lv, nobs_lv, tch):
    """GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059); levels from the module's nlevs (ref.init).  Returns nprof."""
    d = np.asfortranarray(del_, dtype=np.float64)
    n = len(lev)
    zi = np.zeros(max(n, 1), np.int32)
    nprof = C.c_int(0)
    _check(lib().ref_xcov_prs_accum(_f(lev), zi, zi, _i(ks), d, n, d.shape[1], np.zeros((1, 1), np.int32, order="F"), _f(p1), _f(p2),
                                    zi.copy(), l3d, bias_lv, *bias_lv.shape, xcov_lv, *xcov_lv.shape, nobs_lv, *nobs_lv.shape,
                                    C.byref(nprof), int(bool(tch))))
    return nprof.value


def ave_nomiss(x, nobs, xregion, yregion):
    """ave_nomiss (m_gritas_grids.f:1797-1826) on one (im, jm) slab"""
    x, nobs = np.asfortranarray(x, dtype=np.float64), np.asfortranarray(nobs, dtype=np.float64)
    return lib().ref_ave_nomiss(x, nobs, x.shape[0], x.shape[1], _i(xregion), _i(yregion))


def score_regions(regbounds=oracle.REGBOUNDS):
    """first / last latitude row of every region as init_ finds them (m_gritas_grids.f:136-138, 163-175; jm from ref.init)"""
    rb = np.asfortranarray(np.array(regbounds, dtype=np.float64).T)
    out = np.zeros((2, rb.shape[1]), np.int32, order="F")
    lib().ref_region_rows(rb.shape[1], rb, out)
    return out


def maskout(im, jm, lat, lon, maskfld, qcx):
    """maskout_ (m_gritas_masks.F90:137-186); qcx int32 inout, returns the number of observations excluded"""
    return lib().ref_maskout(im, jm, len(lat), _f(lat), _f(lon), np.asfortranarray(maskfld, dtype=np.float64), qcx)


def presort(qcexcl, lat, lon, lev, time, kt, ks, kx):
    """gritas.f:500-508 -> 0-based permutation (the sort itself, m_MergeSorts, is a stable merge sort by assumption)"""
    n = len(lat)
    out = np.zeros(max(n, 1), dtype=np.int32)
    lib().ref_presort(n, out, _i(qcexcl), _f(lat), _f(lon), _f(lev), _i(time), _i(kt), _i(ks), _i(kx))
    return out[:n] - 1


def scores2(g: oracle.Grids, regindx):
    """scores2_ (m_gritas_grids.f:1671-1795) on finalized grids with nr = 1 -> collect_stats (nreg, km, l2d + l3d, 3) float32"""
    nreg = regindx.shape[1]
    out = np.zeros((nreg, g.km, g.l2d + g.l3d, 3), np.float32, order="F")
    _check(lib().ref_scores2(out, nreg, g.im, g.jm, g.km, g.l2d, g.l3d, g.nr, np.asfortranarray(regindx, dtype=np.int32),
                             g.bias_2d, g.rtms_2d, g.stdv_2d, g.nobs_2d, g.bias_3d, g.rtms_3d, g.stdv_3d, g.nobs_3d))
    return out
