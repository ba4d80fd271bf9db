"""Full-size files of the BASELINE.json configs through the hand-written oracle and through the reference's own routines translated to C
(oracle/ref.py): flags, raw sums and statistics must be bit-identical.  One-off check (the test suite does the same on small inputs):
    python tools/oracle_vs_ref_fullsize.py
"""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from gritas_b200 import synth
from oracle import oracle, ref
for name, scale, nfiles in (("cfg1", 1.0, 1), ("cfg2", 1.0, 3), ("cfg5", 1.0, 2), ("cfg3_amsua", 0.25, 1)):
    cfg = synth.get_config(name, scale)
    nr = cfg.nr
    table, l2d, l3d = synth.kxkt_table(cfg)
    t0 = time.time()
    p1, p2 = oracle.pint(cfg.plevs)
    ref.init(cfg.im, cfg.jm, cfg.km, cfg.nlevs, cfg.plevs, table.shape[0], table.shape[1])
    r1, r2 = ref.pint(cfg.nlevs)
    assert np.array_equal(p1, r1) and np.array_equal(p2, r2)
    go = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    gr = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    nobs = 0
    for i in range(nfiles):
        c = synth.make_file(cfg, i)
        nobs += len(c["lat"])
        fo = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        fr = ref.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        assert np.array_equal(fo, fr)
        d = np.zeros((len(fo), nr), order="F"); d[:, 0] = c["omf"]
        if nr > 1: d[:, 1] = c["oma"]
        oracle.accum(go, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fo)
        ref.accum(gr, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, r1, r2, fr)
        assert np.array_equal(fo, fr)
    for nm in go.names():
        assert np.array_equal(getattr(go, nm), getattr(gr, nm)), (name, "raw", nm)
    oracle.norm(go, cfg.nlevs, True, 0)
    ref.norm(gr, True, None)
    for nm in go.names():
        if (nm.endswith("2d") and l2d == 0) or (nm.endswith("3d") and l3d == 0): continue
        assert np.array_equal(getattr(go, nm), getattr(gr, nm)), (name, "norm", nm)
    print(f"{name}: {nfiles} file(s), {nobs} obs, grid {cfg.im}x{cfg.jm}x{cfg.km}, l2d={l2d} l3d={l3d} nr={nr}: flags, raw sums and statistics "
          f"bit-identical ({int(go.nobs_3d.sum() + go.nobs_2d.sum())} binned, {time.time()-t0:.1f} s)", flush=True)
