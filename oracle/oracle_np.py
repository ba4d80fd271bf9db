"""Independent numpy twin of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Written separately from oracle/gritas_oracle.c (vectorised, no shared code) so that the two
restatements of the reference pin each other: tests/test_oracle.py requires them to agree
bit-for-bit.  The C oracle is additionally pinned against the reference's own routines translated to C (oracle/ref.py).
Citations are relative to /root/reference/src/Components/gritas/ unless noted.
"""
from __future__ import annotations

import math

import numpy as np

AMISS = 1.0e15  # m_gritas_grids.f:72


def nint(x):
    """Fortran NINT: round half away from zero, exactly (x - trunc(x) is exact in fp64)."""
    x = np.asarray(x, dtype=np.float64)
    t = np.trunc(x)
    f = x - t
    r = t + (f >= 0.5) - (f <= -0.5)
    r = np.where(np.isnan(x), -2147483648.0, r)
    r = np.clip(r, -2147483648.0, 2147483647.0)
    return r.astype(np.int64)


def pint(plevs):
    """m_gritas_grids.f:217-296."""
    p = np.asarray(plevs, dtype=np.float64)
    n = p.size
    if n <= 0:
        raise ValueError("Pint: must have at least 1 vertical levels")
    if n == 1:
        return p.copy(), p - 0.1
    if np.any(p[:-1] <= p[1:]):
        raise ValueError("Pint: pressure levels do not decrease monotonically")
    if p[-1] == 0.0:
        raise ValueError("Pint: cannot handle 0 hPa")
    # scalar libm (math.log / math.exp), as gfortran would call; numpy's SIMD log/exp loops may differ
    # from glibc by an ulp, which would move a level boundary
    lg = [math.log(float(v)) for v in p]
    p1 = np.empty(n)
    p2 = np.empty(n)
    p1[0] = 2000.0
    for k in range(1, n):
        p1[k] = math.exp(0.5 * (lg[k - 1] + lg[k]))   # exp(0.5*(log(p(k-1))+log(p(k))))
    for k in range(n - 1):
        p2[k] = math.exp(0.5 * (lg[k + 1] + lg[k]))   # exp(0.5*(log(p(k+1))+log(p(k))))
    p2[-1] = 0.0
    return p1, p2


def filter_flags(qcexcl, time, lon, ioptn=1, t1=0, t2=-1, lona=-999.0, lonb=999.0, x_passive=7, x_pre_bad=1):
    """GrITAS_App/gritas.f:710-739."""
    qcexcl = np.asarray(qcexcl)
    passive = (qcexcl == x_passive) | (qcexcl == x_pre_bad)
    want = passive & (ioptn > 0)
    tsel = np.ones(len(qcexcl), bool) if t2 < t1 else (t1 <= np.asarray(time)) & (np.asarray(time) <= t2)
    if lona < -180.0 and lonb > 180.0:
        lsel = np.ones(len(qcexcl), bool)
    else:
        lsel = (lona <= np.asarray(lon)) & (np.asarray(lon) <= lonb)
    keep = tsel & lsel & ((qcexcl == 0) | want)
    return np.where(keep, 0, 1).astype(np.int32)


def box_index(im, jm, nlevs, lat, lon, lev, kx, kt, table, p1, p2, ob_flag):
    """m_gritas_grids.f:405-461 -> dict(ok, surface, i, j, k, l, flag_out, miss) with 0-based i,j,k,l."""
    lat, lon, lev = (np.asarray(a, dtype=np.float64) for a in (lat, lon, lev))
    kx, kt = np.asarray(kx, dtype=np.int64), np.asarray(kt, dtype=np.int64)
    kxmax, ktmax = table.shape
    dlon, dlat = 360.0 / im, 180.0 / (jm - 1)
    live = np.asarray(ob_flag) == 0
    inb = (kx >= 1) & (kx <= kxmax) & (kt >= 1) & (kt <= ktmax)
    lv = np.zeros(len(lat), dtype=np.int64)
    lv[inb] = table[kx[inb] - 1, kt[inb] - 1]
    surface = lv < 0
    l = np.abs(lv)
    in_table = l > 0
    flag_out = np.asarray(ob_flag, dtype=np.int32).copy()
    flag_out[live & ~in_table] = 1
    i = 1 + nint((lon - (-180.0)) / dlon)
    i = np.where(i == im + 1, 1, i)
    i = np.where(i == 0, im, i)
    j = 1 + nint((lat - (-90.0)) / dlat)
    i = np.minimum(im, np.maximum(1, i))
    j = np.minimum(jm, np.maximum(1, j))
    # first kk with lev<=p1(kk) and lev>p2(kk)
    p1 = np.asarray(p1, dtype=np.float64)
    p2 = np.asarray(p2, dtype=np.float64)
    hit = (lev[:, None] <= p1[None, :nlevs]) & (lev[:, None] > p2[None, :nlevs])
    k = np.where(hit.any(axis=1), hit.argmax(axis=1), -1)
    ok = live & in_table
    miss = ok & ~surface & (k < 0)
    return dict(ok=ok, surface=surface, i=i - 1, j=j - 1, k=k, l=l - 1, flag_out=flag_out, miss=miss)


def accum(im, jm, km, nlevs, l2d, l3d, lat, lon, lev, kx, kt, del_, table, p1, p2, ob_flag, grids=None):
    """m_gritas_grids.f:307-478 with sequential (input-order) summation via np.add.at.
    Returns (grids dict of F-order arrays, flag_out, first_miss_index or -1)."""
    nobs = len(lat)
    del_ = np.asarray(del_, dtype=np.float64).reshape(nobs, -1, order="F")
    nr = del_.shape[1]
    nx = 0 if nr == 3 else max(1, nr)
    b = box_index(im, jm, nlevs, lat, lon, lev, kx, kt, table, p1, p2, ob_flag)
    first_miss = int(np.argmax(b["miss"])) if b["miss"].any() else -1
    z = lambda *s: np.zeros(s, dtype=np.float64, order="F")
    if grids is None:
        grids = dict(bias_2d=z(im, jm, max(l2d, 1), nr), rtms_2d=z(im, jm, max(l2d, 1), nr),
                     xcov_2d=z(im, jm, max(l2d, 1), nr), nobs_2d=z(im, jm, max(l2d, 1)),
                     bias_3d=z(im, jm, km, max(l3d, 1), nr), rtms_3d=z(im, jm, km, max(l3d, 1), nr),
                     xcov_3d=z(im, jm, km, max(l3d, 1), nr), nobs_3d=z(im, jm, km, max(l3d, 1)))
    upto = np.arange(nobs) < (first_miss if first_miss >= 0 else nobs)  # the reference stops at the miss
    s2 = b["ok"] & b["surface"] & upto
    s3 = b["ok"] & ~b["surface"] & ~b["miss"] & upto
    for sel, tag in ((s2, "2d"), (s3, "3d")):
        if not sel.any():
            continue
        idx = (b["i"][sel], b["j"][sel], b["l"][sel]) if tag == "2d" else \
              (b["i"][sel], b["j"][sel], b["k"][sel], b["l"][sel])
        for ir in range(nr):
            d = del_[sel, ir]
            np.add.at(grids["bias_" + tag][..., ir], idx, d)
            np.add.at(grids["rtms_" + tag][..., ir], idx, d * d)
        if nx > 0:
            np.add.at(grids["xcov_" + tag][..., 0], idx, del_[sel, 0] * del_[sel, nx - 1])
        np.add.at(grids["nobs_" + tag], idx, 1.0)
    return grids, b["flag_out"], first_miss


def norm(grids, nlevs, nr, nrmlz=True, ntresh=0):
    """m_gritas_grids.f:490-656, vectorised; covers ntresh<=1 (no stale-temporary case)."""
    assert ntresh <= 1, "numpy twin does not model the stale tmp of m_gritas_grids.f:642-644"
    nx = max(1, nr)
    out = {k: v.copy(order="F") for k, v in grids.items()}
    with np.errstate(divide="ignore", invalid="ignore"):
        for tag in ("2d", "3d"):
            if tag == "3d" and nlevs < 2:
                for nm in ("stdv_3d",):
                    out.setdefault(nm, np.zeros_like(out["bias_3d"]))
                break
            n = out["nobs_" + tag]
            m = np.trunc(n) if nrmlz else np.ones_like(n)
            some = (n > 0.0) & (n >= ntresh)
            many = (n > 1.0) & (n >= ntresh)
            one = (n == 1.0)
            stdv = np.zeros_like(out["bias_" + tag])
            tmp = []
            for ir in range(nr):
                bias = out["bias_" + tag][..., ir] / m
                t = out["rtms_" + tag][..., ir] / m
                rt = np.sqrt(t)
                t = np.abs(t - bias * bias)
                tmp.append(t)
                out["bias_" + tag][..., ir] = np.where(some, bias, AMISS)
                out["rtms_" + tag][..., ir] = np.where(some, rt, AMISS)
            xt = out["xcov_" + tag][..., 0] / m
            x1 = np.where(some, xt, AMISS)
            xp = xt - out["bias_" + tag][..., 0] * out["bias_" + tag][..., nx - 1]
            out["xcov_" + tag][..., 0] = x1
            for ir in range(nr):
                full = np.sqrt(m * tmp[ir] / (m - 1))
                if tag == "2d":
                    stdv[..., ir] = np.where(many, full, AMISS)
                else:
                    stdv[..., ir] = np.where(many, full, np.where(one, np.sqrt(tmp[ir]), AMISS))
            xfull = m * xp / (m - 1)
            if tag == "2d":
                out["xcov_" + tag][..., nx - 1] = np.where(many, xfull, AMISS)
            else:
                out["xcov_" + tag][..., nx - 1] = np.where(many, xfull, np.where(one, xp, AMISS))
            out["stdv_" + tag] = stdv
    return out


def select_del_all(iopt, omf, oma, obs, xvec, xm, tch=False, self_=False):
    """GrITAS_App/gritas.f:588-704 (+ :738), numpy twin for the single-ODS residual options >= 4."""
    a = abs(iopt)
    omf, oma, obs, xvec, xm = (np.asarray(v, dtype=np.float64) for v in (omf, oma, obs, xvec, xm))
    with np.errstate(all="ignore"):
        cols = {
            4: [obs], 6: [xvec], 8: [xm],
            10: [oma - omf, -omf, -oma] if tch else [omf - oma, omf],
            12: [-oma, omf - oma, omf] if tch else [omf - oma, oma],
            14: ([obs, obs - omf, obs - oma] if self_ else [omf, oma, oma - omf]) if tch else [oma, omf],
            16: [omf, xvec], 18: [oma, xvec], 20: [obs + xm], 22: [xvec], 24: [obs - omf], 28: [obs - omf + xm],
            30: [obs - oma],
            32: [np.where(xvec > 0.0, (omf - oma) / xvec, AMISS), np.where(xvec > 0.0, oma / xvec, AMISS)],
        }[a - (a & 1)]
        if a in (17, 19):
            cols = [cols[0] / cols[1], cols[1]]
    return np.asfortranarray(np.stack(cols, axis=1))
