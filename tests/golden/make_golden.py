"""Generates tests/golden/*.npz: small seeded inputs + the CPU oracle's outputs for them.

The reference ships no golden vectors and cannot be built here (Fortran, un-vendored GMAO_Shared).  These fixtures hold
the answers of the CPU oracle (oracle/gritas_oracle.c), cross-checked before writing against the independent numpy twin
AND -- where /root/reference is present -- against the reference's own GrITAS_Pint_ / GrITAS_Accum_ / GrITAS_Norm_
translated to C statement by statement (oracle/ref.py): every array of every fixture is what the reference's routines
return for these inputs (tests/test_oracle_vs_ref.py::test_golden_inputs_through_the_translated_reference re-checks the
committed files).  They travel to the GPU box, where the CUDA path must reproduce them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from gritas_b200 import synth  # noqa: E402
from oracle import oracle, oracle_np, ref  # noqa: E402

CASES = {
    # name: (config, scale, nfiles, nr, im, jm, kt_list, kx_lists)
    "conv_nr1": ("cfg1", 0.002, 2, 1, 72, 46, None, None),
    "conv_nr2": ("cfg2", 0.002, 2, 2, 72, 46, None, None),
    "mixed_2d3d_nr2": ("cfg1", 0.002, 2, 2, 48, 25, [44, 4, 1, 2, 44], [[120, 130], [220], [181, 187], [181], [130]]),
    "amsua_nr1": ("cfg3_amsua", 0.0002, 2, 1, 144, 73, None, None),
    "clustered_nr1": ("cfg5", 0.003, 2, 1, 72, 46, None, None),
}
COLS = ("lat", "lon", "lev", "time", "kt", "kx", "ks", "qcexcl", "omf", "oma")


def build_case(name):
    cfgname, scale, nfiles, nr, im, jm, ktl, kxl = CASES[name]
    cfg = synth.get_config(cfgname, scale)
    cfg.im, cfg.jm = im, jm
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    if ktl is not None:
        cfg.kt_list, cfg.kx_lists = ktl, kxl
        pairs = [(kt, kx) for kt, kxs in zip(ktl, kxl) for kx in kxs]
        for i, c in enumerate(files):
            rng = np.random.default_rng(9000 + i)
            pick = rng.integers(0, len(pairs), len(c["lat"]))
            off = c["kt"] == 7
            c["kt"] = np.where(off, 7, np.asarray([p[0] for p in pairs], np.int32)[pick]).astype(np.int32)
            c["kx"] = np.asarray([p[1] for p in pairs], np.int32)[pick]
    return cfg, files, nr


def run_oracle(cfg, files, nr):
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    tw = None
    flags = []
    for c in files:
        fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE,
                                 x_pre_bad=synth.X_PRE_BAD)
        d = np.zeros((len(fl), nr), order="F")
        d[:, 0] = c["omf"]
        if nr > 1:
            d[:, 1] = c["oma"]
        fl_np = fl.copy()
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        tw, fo, miss = oracle_np.accum(cfg.im, cfg.jm, cfg.km, cfg.nlevs, l2d, l3d, c["lat"], c["lon"], c["lev"],
                                       c["kx"], c["kt"], d, table, p1, p2, fl_np, tw)
        assert miss < 0 and np.array_equal(fo, fl)
        flags.append(fl)
    raw = {nm: getattr(g, nm).copy() for nm in ("bias_2d", "rtms_2d", "xcov_2d", "nobs_2d", "bias_3d", "rtms_3d",
                                                "xcov_3d", "nobs_3d")}
    for nm in raw:   # the two independent restatements must agree bit for bit before a golden is written
        assert np.array_equal(raw[nm], tw[nm]), nm
    if ref.available():   # the reference's own routines (translated): same flags and raw sums, then the same statistics
        ref.init(cfg.im, cfg.jm, cfg.km, cfg.nlevs, cfg.plevs, table.shape[0], table.shape[1])
        r1, r2 = ref.pint(cfg.nlevs)
        assert np.array_equal(r1, p1) and np.array_equal(r2, p2)
        gr = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
        for c, fo in zip(files, flags):
            fr = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
            d = np.zeros((len(fr), nr), order="F")
            d[:, 0] = c["omf"]
            if nr > 1:
                d[:, 1] = c["oma"]
            ref.accum(gr, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, r1, r2, fr)
            assert np.array_equal(fr, fo)
        for nm in raw:
            assert np.array_equal(raw[nm], getattr(gr, nm)), nm
        ref.norm(gr, True, None)
    oracle.norm(g, cfg.nlevs, True, 0)
    if ref.available():
        for nm in g.names():
            if not ((nm.endswith("2d") and l2d == 0) or (nm.endswith("3d") and l3d == 0)):
                assert np.array_equal(getattr(g, nm), getattr(gr, nm)), nm
    twn = oracle_np.norm(tw, cfg.nlevs, nr, True, 0)
    for nm in g.names():
        if (nm.endswith("2d") and l2d == 0) or (nm.endswith("3d") and l3d == 0):
            continue                                  # dummy planes of an absent grid family
        assert np.array_equal(getattr(g, nm), twn[nm], equal_nan=True), nm
    return g, raw, flags, (table, l2d, l3d, p1, p2)


def pack(arr, fill):
    """Sparse form of a mostly-`fill` array: flat F-order indices + values."""
    flat = arr.ravel(order="F")
    idx = np.nonzero(flat != fill)[0]
    return idx.astype(np.int64), flat[idx]


def main():
    for name in CASES:
        cfg, files, nr = build_case(name)
        g, raw, flags, (table, l2d, l3d, p1, p2) = run_oracle(cfg, files, nr)
        out = dict(im=cfg.im, jm=cfg.jm, km=cfg.km, nr=nr, l2d=l2d, l3d=l3d, plevs=cfg.plevs, p1=p1, p2=p2,
                   kt_list=np.asarray(cfg.kt_list, np.int32), nfiles=len(files),
                   kx_flat=np.asarray([k for ks in cfg.kx_lists for k in ks], np.int32),
                   kx_len=np.asarray([len(ks) for ks in cfg.kx_lists], np.int32),
                   kxmax=cfg.kxmax, ktmax=cfg.ktmax)
        for i, c in enumerate(files):
            for k in COLS:
                out[f"f{i}_{k}"] = c[k]
            out[f"f{i}_flag"] = flags[i]
        for nm in g.names():
            fill = 0.0 if nm.startswith("nobs") else oracle.AMISS
            out[f"norm_{nm}_idx"], out[f"norm_{nm}_val"] = pack(getattr(g, nm), fill)
        for nm, a in raw.items():
            out[f"raw_{nm}_idx"], out[f"raw_{nm}_val"] = pack(a, 0.0)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", int(g.nobs_3d.sum() + g.nobs_2d.sum()), "obs binned")


if __name__ == "__main__":
    main()
