"""Seeded synthetic ODS-shaped observation columns for the five BASELINE.json configs.

There is no sample ODS data in the reference and no NetCDF reader in this image, so the columns of
one synoptic time are generated directly in the in-memory struct-of-arrays form that
`ods%data%{lat,lon,lev,time,kt,kx,ks,obs,omf,oma,xm,xvec,qcexcl}` has after `ODS_Get`
(reference src/Applications/GrITAS_App/gritas.f:427, 510-551): fp64 reals (FREAL8 build), int32
integers.  Distributions follow SURVEY.md section 8(d).  Everything is a pure function of
(config name, file index), seed = 20241019 + file_index.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

SEED0 = 20241019
X_PASSIVE = 7   # PSAS passive (m_odsmeta, not in tree; hint at grisas.f:600-606)
X_PRE_BAD = 1   # GSI passive (m_odsmeta, not in tree)

# 26 descending pressure levels 1000..10 hPa (subset of the 42 in
# src/Components/gritas/GIO/ctl_files/merra.6hourly3d_omf_p.ctl:8-10)
PLEVS26 = np.array([1000, 975, 950, 925, 900, 850, 800, 750, 700, 650, 600, 550, 500, 450, 400, 350, 300,
                    250, 200, 150, 100, 70, 50, 30, 20, 10], dtype=np.float64)

# (kt, var, kx) rows of src/Components/gritas/etc/gritas_upconv.rc:19-34
UPCONV_ROWS = [(44, "vtmpraob", 120), (44, "tacairep", 130), (44, "tacsdar", 131), (44, "tdropsnd", 132),
               (44, "tacmdcars", 133), (4, "uwndraob", 220), (5, "vwndraob", 220), (4, "uacairep", 230),
               (5, "vacairep", 230), (4, "uacsdar", 231), (5, "vacsdar", 231), (4, "udropsnd", 232),
               (5, "vdropsnd", 232), (4, "uacmdcars", 233), (5, "vacmdcars", 233), (11, "qraob", 120)]


@dataclass
class Config:
    name: str
    im: int
    jm: int
    plevs: np.ndarray            # descending (gritas.f:1819 reverses ascending lists)
    kt_list: list                # one entry per rc row
    kx_lists: list               # kx's of each rc row
    nfiles: int
    nobs_per_file: int
    nr: int                      # 1 = omf ; 2 = omf+oma in one pass
    kind: str = "conv"           # conv | radiance | stratified | clustered
    kxmax: int = 1000            # m_odsmeta::KXMAX is not in tree; rc files use kx <= 960
    ktmax: int = 128             # m_odsmeta::KTMAX is not in tree; rc files use kt <= 89
    kt_surf: tuple = (1, 2, 3)   # gritas_kxkt.f:36-40 ("kt = 1, 2, 3 ... surface")
    extra: dict = field(default_factory=dict)

    @property
    def nlevs(self):
        return len(self.plevs)

    @property
    def km(self):
        return len(self.plevs)  # gritas.f:1885


def _conv_lists():
    return [r[0] for r in UPCONV_ROWS], [[r[2]] for r in UPCONV_ROWS]


def get_config(name: str, scale: float = 1.0) -> Config:
    """BASELINE.json configs.  `scale` shrinks nobs_per_file (parity tests use small scales)."""
    kt, kx = _conv_lists()
    n = lambda x: max(1, int(round(x * scale)))
    if name == "cfg1":   # one synoptic time, ~1M obs, omf, 1x1 deg x 26 levels
        return Config("cfg1", 360, 181, PLEVS26, kt, kx, 1, n(1_000_000), 1)
    if name == "cfg2":   # one month of 6-hourly files, omf+oma
        return Config("cfg2", 360, 181, PLEVS26, kt, kx, 124, n(1_210_000), 2)
    if name == "cfg3_amsua":  # 0.25 deg, AMSU-A channels as levels (etc/gritas_amsua_n15.rc:8-17)
        return Config("cfg3_amsua", 1440, 721, np.arange(15, 0, -1, dtype=np.float64), [40], [[315]],
                      124, n(16_100_000), 1, kind="radiance")
    if name == "cfg3_iasi":   # 0.25 deg, IASI 616 channels (etc/gritas_iasi_metop-b.rc:10-25)
        return Config("cfg3_iasi", 1440, 721, np.arange(616, 0, -1, dtype=np.float64), [40], [[876]],
                      124, n(16_100_000), 1, kind="radiance")
    if name == "cfg4":   # hundreds of single-(kt,kx) grids; exceeds gritas.h l3d_max=22
        ngr = 300
        return Config("cfg4", 360, 181, PLEVS26, [44] * ngr, [[100 + g] for g in range(ngr)],
                      124, n(1_210_000), 2, kind="stratified", extra={"ngrids": ngr})
    if name == "cfg5":   # radiosonde-site / swath hot spots, deterministic mode
        return Config("cfg5", 360, 181, PLEVS26, kt, kx, 124, n(1_210_000), 1, kind="clustered",
                      extra={"nsites": 800})
    raise KeyError(name)


def kxkt_table(cfg: Config):
    """GrITAS_KxKt (reference gritas_kxkt.f:66-108) restated for config setup: returns
    (table[kxmax,ktmax] int32 F-order, l2d, l3d).  +l = upper-air grid l, -l = surface grid l."""
    table = np.zeros((cfg.kxmax, cfg.ktmax), dtype=np.int32, order="F")
    l2d = l3d = 0
    for kt, kxs in zip(cfg.kt_list, cfg.kx_lists):
        akt = abs(kt)
        if kt < 0 or akt in cfg.kt_surf:
            l2d += 1
            l = -l2d
        else:
            l3d += 1
            l = l3d
        for kx in kxs:
            if kx < 0:
                break
            table[kx - 1, akt - 1] = l
    return table, l2d, l3d


def _latlon_uniform(rng, n):
    lat = np.degrees(np.arcsin(rng.uniform(-1.0, 1.0, n)))
    lon = rng.uniform(-180.0, 180.0, n)
    return lat, lon


def _qc(rng, n):
    u = rng.random(n)
    other = rng.choice(np.array([2, 3, 8, 10, 31], dtype=np.int32), n)
    return np.where(u < 0.85, 0, np.where(u < 0.90, X_PASSIVE, other)).astype(np.int32)


def _swath_latlon(rng, nfoot, file_index, width_deg=20.0):
    """Footprints along sun-synchronous-like tracks: ~3.5 orbits per 6 h, inclination 98.7 deg."""
    t = np.sort(rng.uniform(0.0, 1.0, nfoot))                      # fraction of the 6-h window
    phase = 2.0 * np.pi * (3.55 * t + 0.37 * file_index)
    inc = np.radians(98.7)
    lat_c = np.degrees(np.arcsin(np.sin(inc) * np.sin(phase)))
    lon_c = np.degrees(np.arctan2(np.cos(inc) * np.sin(phase), np.cos(phase))) - 360.0 * 0.25 * (t + file_index)
    lat = np.clip(lat_c + rng.uniform(-2.0, 2.0, nfoot), -90.0, 90.0)
    lon = lon_c + rng.uniform(-width_deg / 2, width_deg / 2, nfoot) / np.maximum(np.cos(np.radians(lat)), 0.2)
    lon = (lon + 180.0) % 360.0 - 180.0
    return lat, lon


def make_file(cfg: Config, file_index: int = 0) -> dict:
    """Columns of one synoptic-time file, in file (unsorted) order."""
    rng = np.random.default_rng(SEED0 + file_index)
    n = cfg.nobs_per_file
    if cfg.kind == "radiance":
        nch = cfg.nlevs
        nfoot = max(1, n // nch)
        n = nfoot * nch
        flat, flon = _swath_latlon(rng, nfoot, file_index)
        lat = np.repeat(flat, nch)
        lon = np.repeat(flon, nch)
        lev = np.tile(np.arange(1, nch + 1, dtype=np.float64), nfoot)
        ks = np.repeat(np.arange(1, nfoot + 1, dtype=np.int32), nch)
        kt = np.full(n, cfg.kt_list[0], dtype=np.int32)
        kx = np.full(n, cfg.kx_lists[0][0], dtype=np.int32)
        omf = rng.normal(0.0, 0.3, n)
        time = np.repeat(rng.integers(-180, 181, nfoot).astype(np.int32), nch)
    else:
        lat, lon = _latlon_uniform(rng, n)
        lev = np.exp(rng.uniform(np.log(10.0), np.log(1050.0), n))
        ks = rng.integers(1, max(2, n // 30), n).astype(np.int32)
        time = rng.integers(-180, 181, n).astype(np.int32)
        if cfg.kind == "stratified":
            ngr = cfg.extra["ngrids"]
            w = 1.0 / np.arange(1, ngr + 1) ** 1.1                # Zipf over the kx ids
            row = rng.choice(ngr, n, p=w / w.sum())
        else:
            row = rng.integers(0, len(cfg.kt_list), n)
        kt = np.asarray(cfg.kt_list, dtype=np.int32)[row]
        kx = np.asarray([k[0] for k in cfg.kx_lists], dtype=np.int32)[row]
        off = rng.random(n) < 0.10                                 # 10 % not in the table
        kt = np.where(off, np.int32(7), kt).astype(np.int32)      # kt=7 (q mixing ratio) is in no row
        omf = rng.normal(0.1, 1.5, n)
        if cfg.kind == "clustered":
            ns = cfg.extra["nsites"]
            srng = np.random.default_rng(SEED0 - 1)               # same sites in every file
            slat, slon = _latlon_uniform(srng, ns)
            at_site = rng.random(n) < 0.70
            site = rng.integers(0, ns, n)
            lat = np.where(at_site, slat[site], lat)
            lon = np.where(at_site, slon[site], lon)
            lev = np.where(at_site, cfg.plevs[rng.integers(0, cfg.nlevs, n)], lev)
            ks = np.where(at_site, site + 1, ks).astype(np.int32)
            nsw = int(0.2 * n)                                     # narrow swaths for 20 %
            wlat, wlon = _swath_latlon(rng, nsw, file_index, width_deg=3.0)
            idx = np.nonzero(~at_site)[0][:nsw]
            lat[idx] = wlat[: idx.size]
            lon[idx] = wlon[: idx.size]
    qcexcl = _qc(rng, n)
    oma = 0.6 * omf + rng.normal(0.0, 0.5, n)
    obs = rng.normal(250.0, 30.0, n)
    xvec = rng.uniform(0.5, 2.0, n)
    xm = rng.normal(0.0, 0.2, n)
    return dict(lat=lat, lon=lon, lev=lev, time=time, kt=kt, kx=kx, ks=ks, qcexcl=qcexcl,
                obs=obs, omf=omf, oma=oma, xvec=xvec, xm=xm)


def write_flat(path, cols) -> None:
    """One synoptic time as the flat column file read by csrc/gritas_driver.cpp (no NetCDF in this image)."""
    n = len(cols["lat"])
    with open(path, "wb") as f:
        np.asarray([n], np.int32).tofile(f)
        for k in ("lat", "lon", "lev"):
            np.ascontiguousarray(cols[k], np.float64).tofile(f)
        for k in ("time", "kt", "kx", "ks", "qcexcl"):
            np.ascontiguousarray(cols[k], np.int32).tofile(f)
        for k in ("obs", "omf", "oma", "xvec"):
            np.ascontiguousarray(cols[k], np.float64).tofile(f)


def write_setup(path, cfg: Config) -> None:
    """Grid / level / kt-kx list description for csrc/gritas_driver.cpp."""
    listsz = len(cfg.kt_list)
    kxl = np.full((cfg.kxmax, listsz), -1, dtype=np.int32, order="F")
    for j, ks in enumerate(cfg.kx_lists):
        kxl[: len(ks), j] = np.asarray(ks, np.int32)
    with open(path, "wb") as f:
        np.asarray([cfg.im, cfg.jm, cfg.nlevs, cfg.kxmax, cfg.ktmax, listsz], np.int32).tofile(f)
        np.ascontiguousarray(cfg.plevs, np.float64).tofile(f)
        np.asarray(cfg.kt_list, np.int32).tofile(f)
        kxl.ravel(order="F").tofile(f)
        np.asarray([len(cfg.kt_surf)], np.int32).tofile(f)
        np.asarray(cfg.kt_surf, np.int32).tofile(f)
