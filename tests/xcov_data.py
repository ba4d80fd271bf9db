"""Synthetic radiance profiles for the channel cross-covariance tests: complete profiles of km consecutive observations
(lev = channel number, as in the reference's satellite rc files, etc/gritas_iasi_metop-b.rc), a subset of active channels,
some profiles with a flagged / not-in-table / missing active channel."""
import numpy as np


def make_profiles(seed, nprof, km, achan, nr, kxmax=1000, ktmax=128, two_grids=False):
    rng = np.random.default_rng(seed)
    chans = np.arange(1, km + 1, dtype=np.float64)            # channel numbers 1..km in every profile
    lev = np.tile(chans, nprof)
    nobs = nprof * km
    kt = np.full(nobs, 40, np.int32)
    kx = np.full(nobs, 315, np.int32)
    table = np.zeros((kxmax, ktmax), np.int32, order="F")
    table[315 - 1, 40 - 1] = 1
    l3d = 1
    if two_grids:                                             # a second instrument binned into a second grid
        table[316 - 1, 40 - 1] = 2
        l3d = 2
        second = rng.random(nprof) < 0.4
        kx[np.repeat(second, km)] = 316
    base = rng.normal(0.0, 1.0, (nobs, 1))
    del_ = np.asfortranarray(base * rng.uniform(0.5, 1.5, (1, nr)) + rng.normal(0.0, 0.3, (nobs, nr)) + 0.05)
    flag = np.zeros(nobs, np.int32)
    act = np.isin(np.arange(1, km + 1), achan)
    # rejected observations: on an active channel the whole profile drops out, elsewhere nothing happens
    bad = rng.random(nobs) < 0.01
    flag[bad] = 1
    # not in the table (kx unknown) on some observations
    unk = rng.random(nobs) < 0.005
    kx[unk] = 999
    # a few profiles lose an active channel (lev replaced by a channel number that is not active: m1 /= nchan)
    inactive = np.setdiff1d(np.arange(1, km + 1), achan)
    if len(inactive):
        for p in rng.choice(nprof, max(1, nprof // 20), replace=False):
            k = rng.choice(np.nonzero(act)[0])
            lev[p * km + k] = float(inactive[0])
    return dict(lev=lev, kx=kx, kt=kt, del_=del_, flag=flag, table=table, l3d=l3d)


def reference_numpy(km, achan, d, nr, tch):
    """Independent formulation: dense matrices over the accepted profiles, X^T Y in one go (float64, pairwise order
    differs from the sequential loop: compare with a tolerance)."""
    lev, kx, kt, del_, flag, table = d["lev"], d["kx"], d["kt"], d["del_"], d["flag"], d["table"]
    nobs = len(lev)
    nprof = nobs // km
    nchan = len(achan)
    l3d = d["l3d"]
    bias = np.zeros((nchan, l3d, nr))
    xcov = np.zeros((nchan, nchan, l3d, nr))
    nob = np.zeros((nchan, l3d))
    flag_out = flag.copy()
    acc = 0
    for p in range(nprof):
        sl = slice(p * km, (p + 1) * km)
        c = np.rint(lev[sl]).astype(int)
        m = np.isin(c, achan)
        idx = np.nonzero(m)[0] + p * km
        l = np.abs(table[kx[idx] - 1, kt[idx] - 1])
        flag_out[idx[l <= 0]] = 1
        if len(idx) != nchan or np.any(flag[idx] != 0) or np.any(l <= 0):
            continue
        acc += 1
        r = del_[idx, :]
        for s in range(nr if tch else 1):
            a, b = (r[:, s], r[:, s]) if tch else (r[:, 0], r[:, 1])
            for ll in np.unique(l):
                cols = l == ll
                xcov[:, cols, ll - 1, s] += np.outer(a, b[cols])
        for kk in range(nchan):
            bias[kk, l[kk] - 1, :] += r[kk, :]
            nob[kk, l[kk] - 1] += 1.0
    return bias, xcov, nob, flag_out, acc
