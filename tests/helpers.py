"""Shared helpers of the parity tests: run the same seeded inputs through the CUDA path (via the
C ABI) and through the CPU oracle, and compare in the reference's array layout."""
from __future__ import annotations

import numpy as np

from gritas_b200 import cabi, synth
from gritas_b200.m_gritas_grids import Grids
from oracle import oracle

# fast (tiled/staged accumulate), fast with the direct fp64-RED accumulate, deterministic on the tiled path (sequence
# numbers in the parked entries, k_tile_det), deterministic with the whole-file radix sort + in-order segment sums
MODES = [cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, cabi.MODE_FAST | cabi.FLAG_NO_STAGING, cabi.MODE_DETERMINISTIC,
         cabi.MODE_DETERMINISTIC | cabi.FLAG_DET_SORT]


def is_det(mode):
    return (mode & 0xff) == cabi.MODE_DETERMINISTIC


NAMES = ["bias_2d", "rtms_2d", "stdv_2d", "xcov_2d", "nobs_2d", "bias_3d", "rtms_3d", "stdv_3d", "xcov_3d", "nobs_3d"]


def setup(cfg: synth.Config):
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    return table, l2d, l3d, p1, p2


def del_of(cols, nr):
    n = len(cols["lat"])
    d = np.zeros((n, nr), order="F")
    d[:, 0] = cols["omf"]
    if nr > 1:
        d[:, 1] = cols["oma"]
    if nr > 2:
        d[:, 2] = cols["xm"]
    return d


def oracle_run(cfg, files, nr, flags_in=None, nrmlz=True, ntresh=0, norm=True, km=None):
    """Reference-order run: Accum over the files, then Norm.  Returns (Grids, [flag_out per file])."""
    table, l2d, l3d, p1, p2 = setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, km or cfg.km, l2d, l3d, nr)
    flags = []
    for fi, cols in enumerate(files):
        fl = (np.zeros(len(cols["lat"]), np.int32) if flags_in is None else flags_in[fi].copy())
        oracle.accum(g, cfg.nlevs, cols["lat"], cols["lon"], cols["lev"], cols["kx"], cols["kt"], del_of(cols, nr),
                     table, p1, p2, fl)
        flags.append(fl)
    if norm:
        oracle.norm(g, cfg.nlevs, nrmlz, ntresh)
    return g, flags


def gpu_run(cfg, files, nr, mode=cabi.MODE_FAST, flags_in=None, nrmlz=True, ntresh=None, norm=True, km=None,
            want_flags=True, device_inputs=False):
    """Same through libgritas_b200.so (GrITAS_Accum / GrITAS_Norm mirror).  Returns (dict, flags)."""
    table, l2d, l3d, p1, p2 = setup(cfg)
    G = Grids(cfg.im, cfg.jm, km or cfg.km, cfg.plevs, mode=mode)
    flags = []
    keep = []
    for fi, cols in enumerate(files):
        n = len(cols["lat"])
        fl = None
        if want_flags:
            fl = np.zeros(n, np.int32) if flags_in is None else flags_in[fi].copy()
        d = del_of(cols, nr)
        args = [cols["lat"], cols["lon"], cols["lev"], cols["kx"], cols["kt"], d]
        if device_inputs:
            import torch
            args = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda() for a in args]
            keep.append(args)
        G.GrITAS_Accum(*args, n, nr, table, p1, p2, fl, l2d, l3d=l3d)
        flags.append(fl)
    out = G.alloc_grids()
    if norm:
        G.norm_into(out, nrmlz, ntresh)
    else:
        G.get_raw(out["bias_2d"] if l2d else None, out["rtms_2d"] if l2d else None, out["xcov_2d"] if l2d else None,
                  out["nobs_2d"] if l2d else None, out["bias_3d"] if l3d else None, out["rtms_3d"] if l3d else None,
                  out["xcov_3d"] if l3d else None, out["nobs_3d"] if l3d else None)
    G.close()
    return out, flags


def rel_err(got, ref, scale):
    """|got-ref| / max(|ref|, scale) over finite, non-missing boxes (BASELINE.md parity bar)."""
    ok = np.isfinite(ref) & (np.abs(ref) < 1e14)
    if not ok.any():
        return 0.0
    den = np.maximum(np.abs(ref[ok]), np.broadcast_to(scale, ref.shape)[ok])
    den = np.where(den == 0.0, 1.0, den)
    return float(np.max(np.abs(got[ok] - ref[ok]) / den))


def _fin(x):
    return np.where(np.isfinite(x) & (np.abs(x) < 1e14), np.abs(x), 0.0)


def compare(out, ref: "oracle.Grids", l2d, l3d, nr, exact: bool, tol: float, raw=False):
    """Counts bit-exact always; statistics bit-exact (deterministic mode) or within `tol` relative to
    max(|ref|, natural scale of the box) (fast mode; the mean cancels, so its scale is the rms)."""
    nx = max(1, nr)
    for tag, on in (("2d", l2d), ("3d", l3d)):
        if not on:
            continue
        n_ref = getattr(ref, "nobs_" + tag)
        assert np.array_equal(out["nobs_" + tag], n_ref), f"nobs_{tag} not bit-exact"
        rt = getattr(ref, "rtms_" + tag)                       # (..., nr): rms, or Sxx when raw
        if raw:
            sxx = _fin(rt)
            scales = {"bias": np.sqrt(sxx * np.maximum(n_ref[..., None], 1.0)), "rtms": sxx,
                      "xcov": np.sqrt(sxx[..., 0] * sxx[..., nx - 1])[..., None]}
        else:
            rms = _fin(rt)
            scales = {"bias": rms, "rtms": rms, "stdv": rms,
                      "xcov": (rms[..., 0] * rms[..., nx - 1])[..., None]}
        for nm, scale in scales.items():
            a, b = out[f"{nm}_{tag}"], getattr(ref, f"{nm}_{tag}")
            if raw and nm == "xcov":
                a, b = a[..., :1], b[..., :1]
            assert np.array_equal(np.isnan(a), np.isnan(b)), f"{nm}_{tag} nan pattern"
            assert np.array_equal(a == oracle.AMISS, b == oracle.AMISS), f"{nm}_{tag} AMISS pattern"
            assert np.array_equal(np.isinf(a), np.isinf(b)), f"{nm}_{tag} inf pattern"
            if exact:
                assert np.array_equal(a, b, equal_nan=True), f"{nm}_{tag} not bit-exact"
            else:
                e = rel_err(a, b, np.broadcast_to(scale, b.shape))
                assert e <= tol, f"{nm}_{tag}: rel err {e:.3e} > {tol:.1e}"


# ---- golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py) ---------------------
import glob
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GCOLS = ("lat", "lon", "lev", "time", "kt", "kx", "ks", "qcexcl", "omf", "oma")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = synth.Config(name, int(z["im"]), int(z["jm"]), z["plevs"], [int(v) for v in z["kt_list"]], [], int(z["nfiles"]),
                       0, int(z["nr"]), kxmax=int(z["kxmax"]), ktmax=int(z["ktmax"]))
    pos = 0
    for ln in z["kx_len"]:
        cfg.kx_lists.append([int(v) for v in z["kx_flat"][pos:pos + int(ln)]])
        pos += int(ln)
    files = [{k: z[f"f{i}_{k}"] for k in GCOLS} for i in range(cfg.nfiles)]
    flags = [z[f"f{i}_flag"] for i in range(cfg.nfiles)]

    def dense(prefix, nm, shape, fill):
        a = np.full(int(np.prod(shape)), fill, dtype=np.float64)
        a[z[f"{prefix}_{nm}_idx"]] = z[f"{prefix}_{nm}_val"]
        return a.reshape(shape, order="F")
    l2, l3, nr = max(int(z["l2d"]), 1), max(int(z["l3d"]), 1), cfg.nr
    s2, s3 = (cfg.im, cfg.jm, l2), (cfg.im, cfg.jm, cfg.km, l3)
    shapes = {"bias_2d": s2 + (nr,), "rtms_2d": s2 + (nr,), "stdv_2d": s2 + (nr,), "xcov_2d": s2 + (nr,), "nobs_2d": s2,
              "bias_3d": s3 + (nr,), "rtms_3d": s3 + (nr,), "stdv_3d": s3 + (nr,), "xcov_3d": s3 + (nr,), "nobs_3d": s3}
    norm = {nm: dense("norm", nm, sh, 0.0 if nm.startswith("nobs") else oracle.AMISS) for nm, sh in shapes.items()}
    raw = {nm: dense("raw", nm, sh, 0.0) for nm, sh in shapes.items() if not nm.startswith("stdv")}
    return cfg, files, flags, norm, raw, int(z["l2d"]), int(z["l3d"]), z["p1"], z["p2"]


def filter_in(c, ioptn=1):
    return oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=ioptn, x_passive=synth.X_PASSIVE,
                               x_pre_bad=synth.X_PRE_BAD)
