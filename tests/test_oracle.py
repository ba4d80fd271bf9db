"""The CPU oracle (oracle/gritas_oracle.c) pinned against: the committed golden fixtures, an
independently written numpy twin (bit for bit), and the known-answer tests that follow from the
reference source.  gritas.x itself cannot be built here (no Fortran compiler, un-vendored
GMAO_Shared) and ships no tests; tests/test_oracle_vs_ref.py pins the oracle against the reference's own routines
translated to C statement by statement."""
import numpy as np
import pytest

from gritas_b200 import synth
from oracle import oracle, oracle_np
from tests import helpers as H


@pytest.mark.parametrize("name", H.GOLDEN)
def test_oracle_reproduces_golden(name):
    cfg, files, flags, norm, raw, l2d, l3d, p1, p2 = H.load_golden(name)
    table, a, b = synth.kxkt_table(cfg)
    assert (a, b) == (l2d, l3d)
    o1, o2 = oracle.pint(cfg.plevs)
    assert np.array_equal(o1, p1) and np.array_equal(o2, p2)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, cfg.nr)
    for c, fl_gold in zip(files, flags):
        fl = H.filter_in(c)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, cfg.nr), table, p1, p2, fl)
        assert np.array_equal(fl, fl_gold)
    for nm, a in raw.items():
        assert np.array_equal(getattr(g, nm), a), nm
    oracle.norm(g, cfg.nlevs, True, 0)
    for nm, a in norm.items():
        if (nm.endswith("2d") and not l2d) or (nm.endswith("3d") and not l3d):
            continue
        assert np.array_equal(getattr(g, nm), a), nm


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("nr", [1, 2, 3])
def test_c_oracle_equals_numpy_twin(seed, nr):
    """Two restatements written separately (C loops vs vectorised numpy) agree bit for bit."""
    rng = np.random.default_rng(seed)
    im, jm = int(rng.integers(4, 40)), int(rng.integers(3, 30))
    nlev = int(rng.integers(2, 12))
    plevs = np.sort(rng.uniform(1.0, 1050.0, nlev))[::-1].copy()
    p1, p2 = oracle.pint(plevs)
    kt_list = [44, 4, 1, 2, 5]
    kx_lists = [[10, 11], [12], [10], [13, 14], [12]]
    table, _, _, l2d, l3d = oracle.kxkt(60, 40, kt_list, kx_lists)
    n = 3000
    lat = rng.uniform(-95, 95, n)
    lon = rng.uniform(-200, 380, n)
    lev = np.exp(rng.uniform(np.log(0.5), np.log(1999.0), n))
    kx = rng.integers(8, 17, n).astype(np.int32)
    kt = rng.choice(np.array([44, 4, 1, 2, 5, 7, 0, 61], np.int32), n)
    d = np.asfortranarray(rng.normal(0.3, 2.0, (n, nr)))
    fl0 = (rng.random(n) < 0.2).astype(np.int32)
    # grid-aligned values exercise NINT ties
    lon[:200] = -180.0 + (rng.integers(0, 2 * im, 200) + 0.5) * (360.0 / im)
    lat[:200] = -90.0 + (rng.integers(0, 2 * jm, 200) * 0.5) * (180.0 / (jm - 1))
    lev[200:230] = p2[rng.integers(0, nlev - 1, 30)]
    g = oracle.Grids(im, jm, nlev, l2d, l3d, nr)
    fl = fl0.copy()
    oracle.accum(g, nlev, lat, lon, lev, kx, kt, d, table, p1, p2, fl)
    tw, fo, miss = oracle_np.accum(im, jm, nlev, nlev, l2d, l3d, lat, lon, lev, kx, kt, d, table, p1, p2, fl0)
    assert miss < 0 and np.array_equal(fo, fl)
    for nm in ("bias_2d", "rtms_2d", "xcov_2d", "nobs_2d", "bias_3d", "rtms_3d", "xcov_3d", "nobs_3d"):
        assert np.array_equal(getattr(g, nm), tw[nm]), nm
    keys = oracle.box_keys(im, jm, nlev, nlev, lat, lon, lev, kx, kt, table, p1, p2, fl0, l2d, l3d)
    b = oracle_np.box_index(im, jm, nlev, lat, lon, lev, kx, kt, table, p1, p2, fl0)
    k3 = b["i"] + im * (b["j"] + jm * (b["k"] + nlev * b["l"]))
    k2 = -(4 + b["i"] + im * (b["j"] + jm * b["l"]))
    want = np.where(fl0 != 0, -1, np.where(~b["ok"], -2, np.where(b["surface"], k2, np.where(b["miss"], -3, k3))))
    assert np.array_equal(keys, want)
    if nr <= 2:
        with np.errstate(all="ignore"):
            oracle.norm(g, nlev, True, 0)
            twn = oracle_np.norm(tw, nlev, nr, True, 0)
        for nm in g.names():
            assert np.array_equal(getattr(g, nm), twn[nm], equal_nan=True), nm


def test_nint_is_round_half_away_from_zero():
    x = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999999999999994, 2.4999999999999996, 1e9 + 0.5, np.nan, 1e300])
    assert list(oracle_np.nint(x)) == [1, 2, 3, -1, -2, -3, 0, 2, 1000000001, -2147483648, 2147483647]


def _key(cfg, lat, lon, lev, kx=120, kt=44, flag=0):
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    k = oracle.box_keys(cfg.im, cfg.jm, cfg.km, cfg.nlevs, np.array([lat]), np.array([lon]), np.array([lev]),
                        np.array([kx], np.int32), np.array([kt], np.int32), table, p1, p2, np.array([flag], np.int32),
                        l2d, l3d)[0]
    if k < 0:
        return int(k)
    i = k % cfg.im
    j = (k // cfg.im) % cfg.jm
    kk = (k // (cfg.im * cfg.jm)) % cfg.km
    return (int(i) + 1, int(j) + 1, int(kk) + 1)


def test_kats_from_the_reference_source():
    """SURVEY.md section 4 table; citations are m_gritas_grids.f lines."""
    cfg = synth.get_config("cfg1")
    im, jm = cfg.im, cfg.jm
    dlon = 360.0 / im
    p1, p2 = oracle.pint(cfg.plevs)
    assert _key(cfg, 0.0, 180.0, 500.0)[0] == 1                        # :421 i = im+1 -> 1
    assert _key(cfg, 0.0, -180.0 - dlon / 2 - 1e-9, 500.0)[0] == im    # :422 i = 0 -> im
    assert _key(cfg, 0.0, 270.0, 500.0)[0] == im                       # :425 clamped, not wrapped
    assert _key(cfg, 0.0, -180.0 + dlon / 2, 500.0)[0] == 2            # nint(0.5) = 1
    assert _key(cfg, 0.0, -180.0 + 2.5 * dlon, 500.0)[0] == 4          # nint(2.5) = 3, not banker's 2
    assert _key(cfg, 90.0, 0.0, 500.0)[1] == jm and _key(cfg, -90.0, 0.0, 500.0)[1] == 1
    assert _key(cfg, 1e6, 0.0, 500.0)[1] == jm and _key(cfg, -1e6, 0.0, 500.0)[1] == 1   # :426
    assert _key(cfg, 0.0, 0.0, p2[4])[2] == 6                          # lev == p2(5) -> level 6 (.gt.)
    assert _key(cfg, 0.0, 0.0, p1[4])[2] == 5                          # lev == p1(5) -> level 5 (.le.)
    assert _key(cfg, 0.0, 0.0, 2000.0)[2] == 1
    for bad in (2000.0000001, 0.0, -1.0, np.nan):
        assert _key(cfg, 0.0, 0.0, bad) == -3                          # :458-459 fatal
    assert _key(cfg, 0.0, 0.0, np.nan, kt=7) == -2                     # not in table: never reaches the level search
    assert _key(cfg, 0.0, 0.0, np.nan, flag=1) == -1                   # rejected on entry


def test_filter_c_equals_numpy():
    cfg = synth.get_config("cfg1", 0.01)
    c = synth.make_file(cfg, 0)
    for ioptn in (1, -1, 3, -3):
        for (t1, t2, la, lb) in ((0, -1, -999.0, 999.0), (-60, 60, -999.0, 999.0), (0, -1, -50.0, 70.0), (-180, 0, 10.0, 20.0)):
            a = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn, t1, t2, la, lb, synth.X_PASSIVE, synth.X_PRE_BAD)
            b = oracle_np.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn, t1, t2, la, lb, synth.X_PASSIVE, synth.X_PRE_BAD)
            assert np.array_equal(a, b)
    keep = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], 1, x_passive=7, x_pre_bad=1) == 0
    drop = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], -1, x_passive=7, x_pre_bad=1) == 0
    assert np.array_equal(keep, np.isin(c["qcexcl"], [0, 7, 1])) and np.array_equal(drop, c["qcexcl"] == 0)


def test_presort_is_the_eight_key_stable_order():
    """gritas.f:499-508: stable sorts by qcexcl, lat, lon, lev, time, kt, ks, kx in that order ->
    lexicographic by (kx, ks, kt, time, lev, lon, lat, qcexcl), ties in input order."""
    cfg = synth.get_config("cfg5", 0.005)
    c = synth.make_file(cfg, 1)
    p = oracle.presort(c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], c["ks"], c["kx"])
    want = np.lexsort((c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], c["ks"], c["kx"]))
    assert np.array_equal(p, want)


def test_level_miss_stops_the_oracle_at_the_first_offender():
    cfg = synth.get_config("cfg1", 0.001)
    c = synth.make_file(cfg, 0)
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    ok = np.nonzero(c["kt"] != 7)[0]
    c["lev"][ok[10]] = 3000.0
    c["lev"][ok[20]] = -1.0
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 1)
    with pytest.raises(oracle.LevelMiss) as e:
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), table, p1, p2,
                     np.zeros(len(c["lat"]), np.int32))
    assert e.value.n == ok[10] and e.value.lev == 3000.0
    assert g.nobs_3d.sum() == 10


@pytest.mark.parametrize("iopt", [4, 5, 6, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 23, 25, 28, 29, 31, 32, 33])
def test_residual_options_c_equals_numpy(iopt):
    """gritas.f:588-704: the C restatement of every single-ODS residual option against the numpy twin."""
    rng = np.random.default_rng(iopt)
    n = 500
    omf, oma, xm = rng.normal(0, 2, n), rng.normal(0, 1, n), rng.normal(0, 0.3, n)
    obs = rng.normal(250, 30, n)
    xvec = rng.uniform(-0.5, 2.0, n)          # some non-positive sigo values exercise the `where` of option 32/33
    for tch, self_ in ((False, False), (True, False), (True, True)):
        a = oracle.select_del_all(iopt, omf, oma, obs, xvec, xm, tch=tch, self_=self_)
        b = oracle_np.select_del_all(iopt, omf, oma, obs, xvec, xm, tch=tch, self_=self_)
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    with pytest.raises(ValueError):
        oracle.select_del_all(26, omf, oma, obs, xvec, xm)


def test_norm_3ch_known_answers():
    """GrITAS_3CH_Norm_ (m_gritas_grids.f:1178-1341) on hand-computed boxes: variances (no square root), the 3CH
    estimates with the different slot order of the 2-D and the 3-D block, missing values taking part in the
    whole-array expressions, the n = 1 asymmetry, untouched levels above nlevs, and the early return for nlevs < 2."""
    im, jm, km, nlevs = 2, 1, 3, 2
    g = oracle.Grids(im, jm, km, 1, 1, 3)
    # series with exactly representable statistics: x = {1, 3}, y = {2, 6}, z = {0, 4}
    xs = np.array([[1.0, 3.0], [2.0, 6.0], [0.0, 4.0]])
    for arrs in ((g.nobs_2d, g.bias_2d, g.rtms_2d, (0, 0, 0)), (g.nobs_3d, g.bias_3d, g.rtms_3d, (0, 0, 0, 0))):
        nobs, bias, rtms, ix = arrs
        nobs[ix] = 2.0
        for ir in range(3):
            bias[ix + (ir,)] = xs[ir].sum()
            rtms[ix + (ir,)] = (xs[ir] ** 2).sum()
    # second box: a single observation (2-D: missing variance, 3-D: tmp = |x^2 - x^2| = 0)
    for nobs, bias, rtms, ix in ((g.nobs_2d, g.bias_2d, g.rtms_2d, (1, 0, 0)), (g.nobs_3d, g.bias_3d, g.rtms_3d, (1, 0, 0, 0))):
        nobs[ix] = 1.0
        for ir in range(3):
            bias[ix + (ir,)] = 2.0 + ir
            rtms[ix + (ir,)] = (2.0 + ir) ** 2
    g.bias_3d[0, 0, 2, 0, :] = 7.0          # level 3 > nlevs: never touched
    oracle.norm_3ch(g, nlevs, True, 0)
    v = np.array([2.0, 8.0, 8.0])           # sample variances n/(n-1) * (mean(x^2) - mean(x)^2)
    assert np.array_equal(g.bias_2d[0, 0, 0], [2.0, 4.0, 2.0]) and np.array_equal(g.bias_3d[0, 0, 0, 0], [2.0, 4.0, 2.0])
    assert np.allclose(g.rtms_2d[0, 0, 0], np.sqrt([5.0, 20.0, 8.0]), rtol=0, atol=0)
    assert np.array_equal(g.stdv_2d[0, 0, 0], v) and np.array_equal(g.stdv_3d[0, 0, 0, 0], v)
    e1, ea, eb = 0.5 * (v[0] + v[1] - v[2]), 0.5 * (v[1] + v[2] - v[0]), 0.5 * (v[2] + v[0] - v[1])
    assert np.array_equal(g.xcov_2d[0, 0, 0], [e1, ea, eb])            # :1287-1289
    assert np.array_equal(g.xcov_3d[0, 0, 0, 0], [e1, eb, ea])         # :1337-1339: slots 2 and 3 swapped
    # n = 1
    assert np.all(g.stdv_2d[1, 0, 0] == oracle.AMISS) and np.all(g.xcov_2d[1, 0, 0] == 0.5 * oracle.AMISS)
    assert np.all(g.stdv_3d[1, 0, 0, 0] == 0.0) and np.all(g.xcov_3d[1, 0, 0, 0] == 0.0)
    # empty box of level 2: everything missing, 3CH of three missing values
    assert np.all(g.bias_3d[0, 0, 1, 0] == oracle.AMISS) and np.all(g.xcov_3d[0, 0, 1, 0] == 0.5 * oracle.AMISS)
    # level 3 (> nlevs): untouched sums, variances stay 0, 3CH of zeros
    assert np.all(g.bias_3d[0, 0, 2, 0] == 7.0) and np.all(g.stdv_3d[:, :, 2] == 0.0) and np.all(g.xcov_3d[:, :, 2] == 0.0)
    # nlevs < 2: the 3-D block is never reached (:1292)
    h = oracle.Grids(im, jm, km, 1, 1, 3)
    h.nobs_3d[0, 0, 0, 0] = 2.0
    h.bias_3d[0, 0, 0, 0, :] = 4.0
    oracle.norm_3ch(h, 1, True, 0)
    assert np.all(h.bias_3d[0, 0, 0, 0] == 4.0) and np.all(h.xcov_3d == 0.0) and np.all(h.stdv_3d == 0.0)
    # nr != 3 is rejected
    with pytest.raises(ValueError):
        oracle.norm_3ch(oracle.Grids(im, jm, km, 1, 1, 2), nlevs, True, 0)
