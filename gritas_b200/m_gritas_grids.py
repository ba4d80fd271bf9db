"""Host-side mirror of the reference's `m_gritas_grids` interface for the gridding hot path.

Same names, argument order, argument meaning and error behaviour as the Fortran module
(reference src/Components/gritas/m_gritas_grids.f:85-101) and GrITAS_KxKt
(src/Components/gritas/gritas_kxkt.f:10-12), sitting directly above the C ABI of
include/gritas_b200.h -- it is what fortran/m_gritas_b200.F90 is, written in Python because
this image has no Fortran compiler.  All arithmetic happens in libgritas_b200.so on the GPU.

Differences a caller sees:
  * the grids passed to GrITAS_Accum are NOT touched by it -- the accumulators stay in HBM and
    the host arrays are filled by GrITAS_Norm (or get_raw) once, after the last file;
  * gritas_perror (print + exit(7), gritas_etc.f:54-60) becomes the exception GritasPerror(code=7).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import cabi

AMISS = 1.0e15        # m_gritas_grids.f:72
LonMin = -180.0       # m_gritas_grids.f:33
LatMin = -90.0        # m_gritas_grids.f:42


class GritasPerror(RuntimeError):
    """gritas_perror(msg): the reference prints msg and calls exit(7) (gritas_etc.f:54-60)."""

    def __init__(self, msg, code=7, lev=None):
        super().__init__(msg)
        self.code = code
        self.lev = lev


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def GrITAS_KxKt(KTMAX, KXMAX, L2D_MAX, L3D_MAX, listsz, kt_list, kx_list, ktSurfAll=(1, 2, 3)):
    """gritas_kxkt.f:10-119.  kx_list: (KXMAX, listsz) F-order int32 with a negative terminator per
    column, or a sequence of per-row kx sequences.  Returns (kxkt_table[KXMAX,KTMAX], list_2d, list_3d, l2d, l3d)."""
    L = cabi.load()
    if not isinstance(kx_list, np.ndarray):
        rows = list(kx_list)
        kx_arr = np.full((KXMAX, max(listsz, 1)), -1, dtype=np.int32, order="F")
        for j, r in enumerate(rows[:listsz]):
            kx_arr[: len(r), j] = np.asarray(r, dtype=np.int32)
        kx_list = kx_arr
    kx_list = np.asfortranarray(kx_list, dtype=np.int32)
    kt_list = _i32(kt_list) if listsz else np.zeros(1, np.int32)
    surf = _i32(ktSurfAll) if len(ktSurfAll) else np.zeros(1, np.int32)
    table = np.zeros((KXMAX, KTMAX), dtype=np.int32, order="F")
    l2 = np.zeros(max(L2D_MAX, 1), dtype=np.int32)
    l3 = np.zeros(max(L3D_MAX, 1), dtype=np.int32)
    a, b = C.c_int(0), C.c_int(0)
    rc = L.gritas_b200_kxkt(KTMAX, KXMAX, L2D_MAX, L3D_MAX, listsz, kt_list.ctypes.data, kx_list.ctypes.data,
                            surf.ctypes.data, len(ktSurfAll), table.ctypes.data, l2.ctypes.data, l3.ctypes.data,
                            C.byref(a), C.byref(b))
    if rc:
        raise GritasPerror({1: "KxKt: invalid kt - too large", 2: "KxKt: invalid kt - negative",
                            3: "KxKt: invalid l2d - too large", 4: "KxKt: invalid l3d - too large",
                            5: "KxKt: invalid kx - too large"}[rc])
    return table, l2[: a.value].copy(), l3[: b.value].copy(), a.value, b.value


class Grids:
    """The module state of m_gritas_grids (im, jm, km, nlevs, plevs, dLon, dLat; :24-72) plus the
    device context that replaces the host accumulators."""

    def __init__(self, im, jm, km, plevs, mode=cabi.MODE_FAST, device=-1):
        """grids_init (m_gritas_grids.f:105-190).  plevs descending (gritas.f:1819 reverses them)."""
        self.im, self.jm, self.km = int(im), int(jm), int(km)
        self.plevs = _f64(plevs).copy()
        self.nlevs = int(self.plevs.size)
        self.dLon = 360.0 / self.im          # :128
        self.dLat = 180.0 / (self.jm - 1)    # :129
        self.mode = mode
        self.device = device
        self._h = None
        self._sig = None

    # ---- GrITAS_Pint_ (m_gritas_grids.f:217-296) -------------------------------------------
    def GrITAS_Pint(self):
        L = cabi.load()
        n = self.nlevs
        p1 = np.zeros(max(n, 1))
        p2 = np.zeros(max(n, 1))
        rc = L.gritas_b200_pint(n, self.plevs.ctypes.data if n else None, p1.ctypes.data, p2.ctypes.data)
        if rc:
            raise GritasPerror({1: "Pint: must have at least 1 vertical levels",
                                2: "Pint: pressure levels do not decrease monotonically",
                                3: "Pint: cannot handle 0 hPa"}[rc])
        return p1[:n], p2[:n]

    # ---- context management ---------------------------------------------------------------
    def _ensure(self, nr, kxkt_table, p1, p2, l2d, l3d):
        L = cabi.load()
        table = np.asfortranarray(kxkt_table, dtype=np.int32)
        p1, p2 = _f64(p1), _f64(p2)
        sig = (nr, l2d, l3d, table.shape, hash(table.tobytes()), hash(p1.tobytes()), hash(p2.tobytes()))
        if self._h is not None and sig == self._sig:
            return
        if self._h is not None:
            raise ValueError("GrITAS_Accum called with tables/shapes that differ from the first call")
        h = C.c_void_p()
        rc = L.gritas_b200_create(C.byref(h), self.im, self.jm, self.km, self.nlevs, l2d, l3d, nr, p1.ctypes.data,
                                  p2.ctypes.data, table.ctypes.data, table.shape[0], table.shape[1], self.mode,
                                  self.device)
        if rc == cabi.ERR_NR:
            raise GritasPerror("Accum: could not handle more then 2 residuals")
        if rc == cabi.ERR_CUDA:
            raise RuntimeError("gritas_b200_create: no usable CUDA device (there is no CPU fallback)")
        if rc:
            raise ValueError(f"gritas_b200_create: rc={rc}")
        self._h, self._sig = h, sig
        self.nr, self.l2d, self.l3d = nr, l2d, l3d

    @property
    def handle(self):
        return self._h

    def _err(self):
        return cabi.load().gritas_b200_last_error(self._h).decode()

    def _check(self, rc, bad=None):
        if rc == cabi.OK:
            return
        if rc == cabi.ERR_LEVEL:
            lev = bad.value if bad is not None else None
            raise GritasPerror(f"Accum: obs level is {lev} hPa\nAccum: could not find level", 7, lev)
        if rc == cabi.ERR_NR:
            raise GritasPerror("Accum: could not handle more then 2 residuals")
        raise RuntimeError(f"gritas_b200 rc={rc}: {self._err()}")

    # ---- GrITAS_Accum_ (m_gritas_grids.f:307-478) ------------------------------------------
    def GrITAS_Accum(self, lat, lon, lev, kx, kt, del_, nobs, nr, kxkt_table, p1, p2, ob_flag,
                     l2d, bias_2d=None, rtms_2d=None, xcov_2d=None, nobs_2d=None,
                     l3d=0, bias_3d=None, rtms_3d=None, xcov_3d=None, nobs_3d=None):
        """Same argument list as the reference.  del_ is (nobs, nr) column-major.  ob_flag is inout
        (int32) or None.  The grid arguments are accepted for signature parity and left untouched."""
        self._ensure(nr, kxkt_table, p1, p2, l2d, l3d)
        L = cabi.load()
        bad = C.c_double(0.0)
        if isinstance(del_, np.ndarray) and not del_.flags.f_contiguous:
            del_ = np.asfortranarray(del_)
        rc = L.gritas_b200_accum(self._h, int(nobs), cabi.ptr(lat, np.float64), cabi.ptr(lon, np.float64),
                                 cabi.ptr(lev, np.float64), cabi.ptr(del_, np.float64), cabi.ptr(kx, np.int32),
                                 cabi.ptr(kt, np.int32), cabi.ptr(ob_flag, np.int32), C.byref(bad))
        self._check(rc, bad)

    def accum_ods(self, ods, iopt=cabi.IOPT_OMF, scale_res=0, ioptn=None, t1=0, t2=-1, lona=-999.0, lonb=999.0,
                  x_passive=7, x_pre_bad=1, kxkt_table=None, p1=None, p2=None, l2d=0, l3d=0, ob_flag_out=None):
        """Fused gritas.f:562-587 (residual selection) + :710-739 (QC/time/lon filter) + GrITAS_Accum_
        straight from the ODS columns (dict with lat lon lev time kt kx qcexcl omf oma [obs xvec])."""
        nr = 2 if abs(iopt) == cabi.IOPT_OMF_OMA else 1
        if self._h is None:
            self._ensure(nr, kxkt_table, p1, p2, l2d, l3d)
        L = cabi.load()
        bad = C.c_double(0.0)
        n = int(ods["lat"].shape[0])
        g = lambda k, dt: cabi.ptr(ods.get(k), dt) if ods.get(k) is not None else None
        rc = L.gritas_b200_accum_ods(self._h, n, g("lat", np.float64), g("lon", np.float64), g("lev", np.float64),
                                     g("time", np.int32), g("kt", np.int32), g("kx", np.int32), g("qcexcl", np.int32),
                                     g("omf", np.float64), g("oma", np.float64), g("obs", np.float64),
                                     g("xvec", np.float64), iopt, scale_res, iopt if ioptn is None else ioptn,
                                     t1, t2, lona, lonb, x_passive, x_pre_bad, cabi.ptr(ob_flag_out, np.int32),
                                     C.byref(bad))
        self._check(rc, bad)

    NR_OF_IOPT = {10: 2, 11: 2, 12: 2, 13: 2, 14: 2, 15: 2, 16: 2, 17: 2, 18: 2, 19: 2, 32: 2, 33: 2, cabi.IOPT_OMF_OMA: 2}

    def accum_odsx(self, ods, iopt, options=0, scale_res=0, ioptn=None, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7,
                   x_pre_bad=1, kxkt_table=None, p1=None, p2=None, l2d=0, l3d=0, ob_flag_out=None):
        """All single-ODS residual options of gritas.f:562-704 fused with the QC filter and GrITAS_Accum_."""
        a = abs(iopt)
        nr = 3 if (options & cabi.OPT_TCH and 10 <= a <= 15) else self.NR_OF_IOPT.get(a, 1)
        if self._h is None:
            self._ensure(nr, kxkt_table, p1, p2, l2d, l3d)
        cols = cabi.OdsCols()
        dt = {"time": np.int32, "kt": np.int32, "kx": np.int32, "qcexcl": np.int32}
        for k, _ in cabi.OdsCols._fields_:
            v = ods.get(k)
            setattr(cols, k, cabi.ptr(v, dt.get(k, np.float64)) if v is not None else None)
        bad = C.c_double(0.0)
        rc = cabi.load().gritas_b200_accum_odsx(self._h, int(ods["lat"].shape[0]), C.byref(cols), iopt, options, scale_res,
                                                iopt if ioptn is None else ioptn, t1, t2, lona, lonb, x_passive, x_pre_bad,
                                                cabi.ptr(ob_flag_out, np.int32), C.byref(bad))
        self._check(rc, bad)

    def accum_odsx_sorted(self, ods, iopt, options=0, scale_res=0, ioptn=None, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7,
                          x_pre_bad=1, ob_flag_out=None, is_out=None):
        """gritas.f:499-523 (pre-sort + gathers) + 562-739 + GrITAS_Accum_ on the device, from unsorted ODS columns (needs ks)."""
        cols = cabi.OdsCols()
        dt = {"time": np.int32, "kt": np.int32, "kx": np.int32, "qcexcl": np.int32}
        for k, _ in cabi.OdsCols._fields_:
            v = ods.get(k)
            setattr(cols, k, cabi.ptr(v, dt.get(k, np.float64)) if v is not None else None)
        bad = C.c_double(0.0)
        rc = cabi.load().gritas_b200_accum_odsx_sorted(self._h, int(ods["lat"].shape[0]), C.byref(cols), cabi.ptr(ods["ks"], np.int32),
                                                       iopt, options, scale_res, iopt if ioptn is None else ioptn, t1, t2, lona, lonb,
                                                       x_passive, x_pre_bad, cabi.ptr(ob_flag_out, np.int32), cabi.ptr(is_out, np.int32),
                                                       C.byref(bad))
        self._check(rc, bad)

    def accum_odsx_meanres(self, ods, mods, iopt, ioptn=None, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7, x_pre_bad=1,
                           ob_flag_out=None):
        """-meanres variants of gritas.f:653-691: `ods` the member, `mods` the ensemble mean (obs, omf, oma)."""
        cols = cabi.OdsCols()
        dt = {"time": np.int32, "kt": np.int32, "kx": np.int32, "qcexcl": np.int32}
        for k, _ in cabi.OdsCols._fields_:
            v = ods.get(k)
            setattr(cols, k, cabi.ptr(v, dt.get(k, np.float64)) if v is not None else None)
        bad = C.c_double(0.0)
        g = lambda k: cabi.ptr(mods.get(k), np.float64) if mods.get(k) is not None else None
        rc = cabi.load().gritas_b200_accum_odsx_meanres(self._h, int(ods["lat"].shape[0]), C.byref(cols), g("obs"), g("omf"), g("oma"),
                                                        iopt, iopt if ioptn is None else ioptn, t1, t2, lona, lonb, x_passive,
                                                        x_pre_bad, cabi.ptr(ob_flag_out, np.int32), C.byref(bad))
        self._check(rc, bad)

    def set_call_order(self, order):
        """Deterministic mode: the place of the next Accum call in the summation order (multi-rank: the global file index)."""
        self._check(cabi.load().gritas_b200_set_call_order(self._h, int(order)))

    def sync(self):
        bad = C.c_double(0.0)
        self._check(cabi.load().gritas_b200_sync(self._h, C.byref(bad)), bad)

    def sync_no_flush(self):
        """Waits for every queued accumulate kernel without touching the tile lists (before a sharded Norm driven from
        one process: the lists of all handles must be complete before any handle pulls from them)."""
        bad = C.c_double(0.0)
        self._check(cabi.load().gritas_b200_wait(self._h, C.byref(bad)), bad)

    def barrier(self):
        self._check(cabi.load().gritas_b200_barrier(self._h))

    def flush(self):
        self._check(cabi.load().gritas_b200_flush(self._h))

    def reset(self):
        """grisas.f:285-296 zeroes the grids at every synoptic time."""
        self._check(cabi.load().gritas_b200_reset(self._h))

    # ---- GrITAS_Norm_ (m_gritas_grids.f:490-656) -------------------------------------------
    def GrITAS_Norm(self, nr, nrmlz, l2d, bias_2d, rtms_2d, stdv_2d, xcov_2d, nobs_2d,
                    l3d, bias_3d, rtms_3d, stdv_3d, xcov_3d, nobs_3d, ntresh=None):
        if nr > 2:
            raise GritasPerror("Accum: could not handle more then 2 residuals")  # :555 (sic)
        if self._h is None:
            raise RuntimeError("GrITAS_Norm before any GrITAS_Accum")
        assert nr == self.nr and l2d == self.l2d and l3d == self.l3d
        p = lambda a: cabi.ptr(a, np.float64)
        rc = cabi.load().gritas_b200_norm(self._h, 1 if nrmlz else 0, -1 if ntresh is None else int(ntresh),
                                          p(bias_2d), p(rtms_2d), p(stdv_2d), p(xcov_2d), p(nobs_2d),
                                          p(bias_3d), p(rtms_3d), p(stdv_3d), p(xcov_3d), p(nobs_3d))
        self._check(rc)

    # ---- pre-sort (gritas.f:499-508) --------------------------------------------------------
    def presort(self, qcexcl, lat, lon, lev, time, kt, ks, kx, index_base=0, out=None):
        """IndexSet + the eight stable IndexSorts of gritas.f:499-508 on the device; returns the permutation `is`
        (int32, 0- or 1-based).  Works before the first GrITAS_Accum (no handle yet)."""
        n = int(lat.shape[0])
        if out is None:
            out = np.zeros(n, np.int32)
        p = cabi.ptr
        rc = cabi.load().gritas_b200_presort(self._h, n, p(qcexcl), p(lat), p(lon), p(lev), p(time), p(kt), p(ks), p(kx),
                                             int(index_base), p(out))
        if rc:
            raise RuntimeError(f"gritas_b200_presort rc={rc}")
        return out

    # ---- GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341; gritas.f:813-818, -tch) -------------
    def GrITAS_3CH_Norm(self, nr, nrmlz, l2d, bias_2d, rtms_2d, var_2d, xcov_2d, nobs_2d,
                        l3d, bias_3d, rtms_3d, var_3d, xcov_3d, nobs_3d, ntresh=None):
        if nr > 3:
            raise GritasPerror("GrITAS_3CH_Norm_ could not handle more then 2 residuals")  # :1240 (sic)
        if self._h is None:
            raise RuntimeError("GrITAS_3CH_Norm before any GrITAS_Accum")
        assert nr == self.nr and l2d == self.l2d and l3d == self.l3d
        p = lambda a: cabi.ptr(a, np.float64)
        rc = cabi.load().gritas_b200_norm_3ch(self._h, 1 if nrmlz else 0, -1 if ntresh is None else int(ntresh),
                                              p(bias_2d), p(rtms_2d), p(var_2d), p(xcov_2d), p(nobs_2d),
                                              p(bias_3d), p(rtms_3d), p(var_3d), p(xcov_3d), p(nobs_3d))
        self._check(rc)

    def get_raw(self, bias_2d, rtms_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, xcov_3d, nobs_3d):
        p = lambda a: cabi.ptr(a, np.float64)
        self._check(cabi.load().gritas_b200_get_raw(self._h, p(bias_2d), p(rtms_2d), p(xcov_2d), p(nobs_2d),
                                                    p(bias_3d), p(rtms_3d), p(xcov_3d), p(nobs_3d)))

    def add_raw(self, bias_2d, rtms_2d, xcov_2d, nobs_2d, bias_3d, rtms_3d, xcov_3d, nobs_3d):
        p = lambda a: cabi.ptr(a, np.float64)
        self._check(cabi.load().gritas_b200_add_raw(self._h, p(bias_2d), p(rtms_2d), p(xcov_2d), p(nobs_2d),
                                                    p(bias_3d), p(rtms_3d), p(xcov_3d), p(nobs_3d)))

    # ---- GrITAS_MinMax_ (m_gritas_grids.f:1019-1167) ---------------------------------------
    def GrITAS_MinMax(self, lat, lon, lev, kx, kt, del_, nobs, nr, kxkt_table, p1, p2, ob_flag, l2d, minv_2d=None,
                      maxv_2d=None, nobs_2d=None, l3d=0, minv_3d=None, maxv_3d=None, nobs_3d=None):
        """Same argument list as the reference; the handle must have been made with FLAG_MINMAX."""
        if nr > 1:
            raise GritasPerror("MinMax: could not handle more then 1 field")
        if not (self.mode & cabi.FLAG_MINMAX):
            raise ValueError("Grids(..., mode=... | cabi.FLAG_MINMAX) needed for GrITAS_MinMax")
        self.GrITAS_Accum(lat, lon, lev, kx, kt, del_, nobs, nr, kxkt_table, p1, p2, ob_flag, l2d, l3d=l3d)

    def get_minmax(self, minv_2d, maxv_2d, nobs_2d, minv_3d, maxv_3d, nobs_3d):
        p = lambda a: cabi.ptr(a, np.float64)
        self._check(cabi.load().gritas_b200_get_minmax(self._h, p(minv_2d), p(maxv_2d), p(nobs_2d), p(minv_3d), p(maxv_3d),
                                                       p(nobs_3d)))

    def set_mask(self, maskfld):
        """gritas_maskout (m_gritas_masks.F90:137-186) fused into the accumulate kernels; (im, jm) fraction field."""
        if maskfld is not None:
            maskfld = np.asfortranarray(maskfld, dtype=np.float64)
            assert maskfld.shape == (self.im, self.jm)
        self._check(cabi.load().gritas_b200_set_mask(self._h, cabi.ptr(maskfld, np.float64)))

    # ---- scorecard (m_gritas_grids.f:1526-1826) --------------------------------------------
    REGNAMES = ("global", "n.hem", "tropics", "s.hem")                       # m_gritas_grids.f:58
    REGBOUNDS = ((-90.0, 90.0), (20.0, 80.0), (-20.0, 20.0), (-80.0, -20.0))  # :146-157

    def score_regions(self, regbounds=None):
        """regindx(2, mregs) as init_ derives it (m_gritas_grids.f:163-175)."""
        rb = np.asfortranarray(np.array(self.REGBOUNDS if regbounds is None else regbounds, dtype=np.float64).T)
        out = np.zeros((2, rb.shape[1]), np.int32, order="F")
        rc = cabi.load().gritas_b200_score_regions(self.jm, rb.shape[1], rb.ctypes.data, out.ctypes.data)
        if rc:
            raise ValueError(f"gritas_b200_score_regions rc={rc}")
        return out

    def GrITAS_Scores(self, regindx=None, collect_stats=None):
        """scores2_ on the current accumulators (Norm with nrmlz = .true. + ave_nomiss per region / level / variable);
        returns collect_stats(mregs, km, l2d + l3d, nevals = 3) real(4).  The handle must have nr = 1 (:1704-1706)."""
        if self.nr != 1:
            raise GritasPerror("GrITAS_scores_: incorrec number of residuals, aborting ...", 98)
        reg = self.score_regions() if regindx is None else np.asfortranarray(regindx, dtype=np.int32)
        if collect_stats is None:
            collect_stats = np.zeros((reg.shape[1], self.km, self.l2d + self.l3d, 3), np.float32, order="F")
        self._check(cabi.load().gritas_b200_scores(self._h, reg.shape[1], reg.ctypes.data, 1, self.im, collect_stats.ctypes.data))
        return collect_stats

    def alloc_grids(self, nr=None, pinned=False):
        """Host arrays shaped like gritas.f:336-348, zero-initialised (gritas.f:352-363)."""
        nr = self.nr if nr is None else nr
        l2, l3 = max(self.l2d, 1), max(self.l3d, 1)
        mk = (lambda *s: cabi.pinned_empty(s)) if pinned else (lambda *s: np.empty(s, dtype=np.float64, order="F"))
        out = {}
        for nm in ("bias", "rtms", "stdv", "xcov"):
            out[nm + "_2d"] = mk(self.im, self.jm, l2, nr)
            out[nm + "_3d"] = mk(self.im, self.jm, self.km, l3, nr)
        out["nobs_2d"] = mk(self.im, self.jm, l2)
        out["nobs_3d"] = mk(self.im, self.jm, self.km, l3)
        for a in out.values():
            a[...] = 0.0
        return out

    def norm_into(self, grids, nrmlz=True, ntresh=None):
        g = grids
        self.GrITAS_Norm(self.nr, nrmlz, self.l2d, g["bias_2d"] if self.l2d else None, g["rtms_2d"] if self.l2d else None,
                         g["stdv_2d"] if self.l2d else None, g["xcov_2d"] if self.l2d else None,
                         g["nobs_2d"] if self.l2d else None, self.l3d, g["bias_3d"] if self.l3d else None,
                         g["rtms_3d"] if self.l3d else None, g["stdv_3d"] if self.l3d else None,
                         g["xcov_3d"] if self.l3d else None, g["nobs_3d"] if self.l3d else None, ntresh)
        return g

    # ---- multi-GPU ------------------------------------------------------------------------
    def comm_init(self, nranks, rank, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._check(cabi.load().gritas_b200_comm_init(self._h, nranks, rank, buf))

    def norm_reduce_into(self, grids, root=0, nrmlz=True, ntresh=None):
        """Collective GrITAS_Norm of a multi-rank run: only the root's arrays are written."""
        g = grids
        p = lambda k, on: cabi.ptr(g[k], np.float64) if on else None
        rc = cabi.load().gritas_b200_norm_reduce(
            self._h, root, 1 if nrmlz else 0, -1 if ntresh is None else int(ntresh),
            p("bias_2d", self.l2d), p("rtms_2d", self.l2d), p("stdv_2d", self.l2d), p("xcov_2d", self.l2d), p("nobs_2d", self.l2d),
            p("bias_3d", self.l3d), p("rtms_3d", self.l3d), p("stdv_3d", self.l3d), p("xcov_3d", self.l3d), p("nobs_3d", self.l3d))
        self._check(rc)
        return g

    # sharded Norm over peer memory (one node): every rank finalizes and writes its slab of rows
    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(cabi.PEER_BLOB_BYTES)
        self._check(cabi.load().gritas_b200_peer_export(self._h, buf))
        return buf.raw

    def peer_attach(self, nranks, rank, blobs):
        """blobs: the peer_export() bytes of every rank, in rank order (all-gathered by the caller)."""
        raw = b"".join(blobs)
        assert len(raw) == nranks * cabi.PEER_BLOB_BYTES
        self._check(cabi.load().gritas_b200_peer_attach(self._h, nranks, rank, C.create_string_buffer(raw, len(raw))))

    def export_planes(self):
        """Phase 1 of the planes exchange for a host that drives several handles without a communicator."""
        self._check(cabi.load().gritas_b200_export_planes(self._h))

    def slab(self, rank, nranks):
        """Rows (jl = j + jm*l, 0-based, half-open) of the 3-D and the 2-D grids that `rank` finalizes and writes."""
        v = [C.c_int(0) for _ in range(4)]
        rc = cabi.load().gritas_b200_slab(self._h, rank, nranks, *[C.byref(x) for x in v])
        if rc:
            raise ValueError(f"gritas_b200_slab rc={rc}")
        return tuple(x.value for x in v)

    def norm_sharded_into(self, grids, nrmlz=True, ntresh=None):
        """Collective GrITAS_Norm of a multi-rank run: this rank's slab of every array is written."""
        g = grids
        p = lambda k, on: cabi.ptr(g[k], np.float64) if on else None
        rc = cabi.load().gritas_b200_norm_sharded(
            self._h, 1 if nrmlz else 0, -1 if ntresh is None else int(ntresh),
            p("bias_2d", self.l2d), p("rtms_2d", self.l2d), p("stdv_2d", self.l2d), p("xcov_2d", self.l2d), p("nobs_2d", self.l2d),
            p("bias_3d", self.l3d), p("rtms_3d", self.l3d), p("stdv_3d", self.l3d), p("xcov_3d", self.l3d), p("nobs_3d", self.l3d))
        self._check(rc)
        return g

    def allreduce(self):
        self._check(cabi.load().gritas_b200_allreduce(self._h))

    # ---- introspection --------------------------------------------------------------------
    def stream(self):
        return cabi.load().gritas_b200_stream(self._h)

    def counters(self):
        L = cabi.load()
        return dict(launches=L.gritas_b200_launch_count(self._h), h2d=L.gritas_b200_h2d_bytes(self._h),
                    d2h=L.gritas_b200_d2h_bytes(self._h))

    def close(self):
        """GrITAS_Clean (gritas.f:2017-2026)."""
        if self._h is not None:
            cabi.load().gritas_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = cabi.load().gritas_b200_nccl_unique_id(buf)
    if rc:
        raise RuntimeError(f"gritas_b200_nccl_unique_id rc={rc}")
    return buf.raw


class XCov:
    """GrITAS_XCov_Accum / GrITAS_XCov_Norm of m_gritas_grids (m_gritas_grids.f:667-1008) on the GPU: channel
    cross-covariances from complete profiles of km observations.  achan: the active channels (the achan(:, ichan_version)
    column); nr = 2 (Desroziers) or nr = 3 with tch (three-cornered hat)."""

    def __init__(self, km, achan, l3d, nr, tch, kxkt_table, device=-1):
        L = cabi.load()
        self.km, self.nchan, self.l3d, self.nr, self.tch = int(km), len(achan), int(l3d), int(nr), bool(tch)
        a = _i32(achan)
        table = np.asfortranarray(kxkt_table, dtype=np.int32)
        h = C.c_void_p()
        rc = L.gritas_b200_xcov_create(C.byref(h), self.km, self.nchan, a.ctypes.data, self.l3d, self.nr, 1 if tch else 0,
                                       table.ctypes.data, table.shape[0], table.shape[1], device)
        self._h = None
        if rc == cabi.ERR_NR:
            raise GritasPerror("GrITAS_XCov_Accum_: improper dim settings / setting inconsistent w/ dims")   # :734-745
        if rc == cabi.ERR_CUDA:
            raise RuntimeError("gritas_b200_xcov_create: no usable CUDA device (there is no CPU fallback)")
        if rc:
            raise ValueError(f"gritas_b200_xcov_create rc={rc}")
        self._h = h

    @classmethod
    def for_levels(cls, p1, p2, l3d, nr, tch, device=-1):
        """Handle for GrITAS_XCov_Prs_Accum (m_gritas_grids.f:1862-2059): level cross-covariances of soundings; p1 / p2 from
        GrITAS_Pint.  Norm / get_raw / reset as for the channel variant (nchan = nlevs)."""
        self = cls.__new__(cls)
        p1, p2 = _f64(p1), _f64(p2)
        self.km = self.nchan = int(p1.size)
        self.l3d, self.nr, self.tch = int(l3d), int(nr), bool(tch)
        self._h = None
        h = C.c_void_p()
        rc = cabi.load().gritas_b200_xcov_prs_create(C.byref(h), self.nchan, p1.ctypes.data, p2.ctypes.data, self.l3d, self.nr,
                                                     1 if tch else 0, device)
        if rc == cabi.ERR_NR:
            raise GritasPerror("GrITAS_XCov_Prs_Accum_: improper dim settings / setting inconsistent w/ dims")   # :1929-1940
        if rc == cabi.ERR_CUDA:
            raise RuntimeError("gritas_b200_xcov_prs_create: no usable CUDA device (there is no CPU fallback)")
        if rc:
            raise ValueError(f"gritas_b200_xcov_prs_create rc={rc}")
        self._h = h
        return self

    def GrITAS_XCov_Prs_Accum(self, lev, kx, kt, ks, del_, nobs, nr, kxkt_table=None, p1=None, p2=None, ob_flag=None, l3d=None,
                              bias_lv=None, xcov_lv=None, nobs_lv=None, tch=None):
        """Reference argument list (m_gritas_grids.f:1862-1865); kx / kt / kxkt_table / ob_flag / the host arrays are accepted and
        ignored, as the reference ignores them.  Returns nprof."""
        assert nr == self.nr
        if isinstance(del_, np.ndarray) and not del_.flags.f_contiguous:
            del_ = np.asfortranarray(del_)
        n, bad = C.c_int(0), C.c_double(0.0)
        rc = cabi.load().gritas_b200_xcov_prs_accum(self._h, int(nobs), cabi.ptr(lev, np.float64), cabi.ptr(ks, np.int32),
                                                    cabi.ptr(del_, np.float64), C.byref(n), C.byref(bad))
        if rc == cabi.ERR_LEVEL:
            raise GritasPerror(f"Accum: could not find level(1) ({bad.value} hPa)", 7, bad.value)
        self._check(rc)
        return n.value

    def _check(self, rc):
        if rc == 0:
            return
        msg = cabi.load().gritas_b200_xcov_last_error(self._h).decode()
        if rc == cabi.ERR_PROFILE:
            raise GritasPerror(msg, 7)
        raise RuntimeError(f"gritas_b200_xcov rc={rc}: {msg}")

    def GrITAS_XCov_Accum(self, lat, lon, lev, kx, kt, del_, nobs, nr, kxkt_table, ob_flag, l3d, bias_lv=None, xcov_lv=None,
                          nobs_lv=None, tch=None):
        """Reference argument list (m_gritas_grids.f:667-670); lat / lon / the host arrays are accepted and ignored (the
        reference never reads lat / lon either; the accumulators live on the device).  Returns nprof."""
        assert nr == self.nr and l3d == self.l3d
        if isinstance(del_, np.ndarray) and not del_.flags.f_contiguous:
            del_ = np.asfortranarray(del_)
        n = C.c_int(0)
        self._check(cabi.load().gritas_b200_xcov_accum(self._h, int(nobs), cabi.ptr(lev, np.float64), cabi.ptr(kx, np.int32),
                                                       cabi.ptr(kt, np.int32), cabi.ptr(del_, np.float64),
                                                       cabi.ptr(ob_flag, np.int32), C.byref(n)))
        return n.value

    def alloc(self):
        return (np.zeros((self.nchan, self.l3d, self.nr), order="F"), np.zeros((self.nchan, self.nchan, self.l3d, self.nr), order="F"),
                np.zeros((self.nchan, self.l3d), order="F"))

    def GrITAS_XCov_Norm(self, nr, nrmlz, tch, self_, l3d, bias_3d, xcov_3d, nobs_3d, ntresh=None):
        """m_gritas_grids.f:859-861: fills the three arrays with the normalised statistics."""
        assert nr == self.nr and bool(tch) == self.tch and l3d == self.l3d
        self._check(cabi.load().gritas_b200_xcov_norm(self._h, 1 if nrmlz else 0, 1 if self_ else 0, -1 if ntresh is None else int(ntresh),
                                                      cabi.ptr(bias_3d, np.float64), cabi.ptr(xcov_3d, np.float64),
                                                      cabi.ptr(nobs_3d, np.float64)))

    def get_raw(self, bias_lv, xcov_lv, nobs_lv):
        self._check(cabi.load().gritas_b200_xcov_get_raw(self._h, cabi.ptr(bias_lv, np.float64), cabi.ptr(xcov_lv, np.float64),
                                                         cabi.ptr(nobs_lv, np.float64)))

    def reset(self):
        self._check(cabi.load().gritas_b200_xcov_reset(self._h))

    def launches(self):
        return cabi.load().gritas_b200_xcov_launch_count(self._h)

    def close(self):
        if self._h is not None:
            cabi.load().gritas_b200_xcov_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
