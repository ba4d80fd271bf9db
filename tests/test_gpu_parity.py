"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): counts and box assignment bit-exact; mean/stdv/rms within 1e-12
relative in the deterministic path (we require BIT-EXACT there: same summation order as the
sequential reference loop) and within 1e-6 in the atomic path.
"""
import numpy as np
import pytest

from gritas_b200 import cabi, synth
from gritas_b200.m_gritas_grids import Grids, GritasPerror
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL_FAST = 1e-6   # north_star tolerance of the atomic path
TOL_DET = 1e-12   # north_star tolerance of the deterministic path (tests demand bit-exact anyway)


def small(name, scale, nfiles, res=(72, 46)):
    """A BASELINE.json config at reduced size: fewer obs per file and the reference's coarsest grid
    (-res a = 72 x 46, gritas.f:1569) so that the oracle's host grids stay small."""
    cfg = synth.get_config(name, scale)
    if res:
        cfg.im, cfg.jm = res
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    return cfg, files


def flags_of(cfg, cols, ioptn=1):
    return oracle.filter_flags(cols["qcexcl"], cols["time"], cols["lon"], ioptn=ioptn)


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("nr", [1, 2])
def test_conventional_two_files(mode, nr):
    cfg, files = small("cfg1", 0.05, 2)
    fin = [flags_of(cfg, c) for c in files]
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, rflags = H.oracle_run(cfg, files, nr, fin)
    out, gflags = H.gpu_run(cfg, files, nr, mode, fin)
    for a, b in zip(gflags, rflags):
        assert np.array_equal(a, b)                      # ob_flag write-back (m_gritas_grids.f:413-416)
    H.compare(out, ref, l2d, l3d, nr, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("mode", H.MODES)
def test_surface_and_upper_air_mix(mode):
    """kt 1,2,3 are surface (2-D grids, lev ignored: m_gritas_grids.f:432-441); same (kx,kt) in two
    rows -> last wins (gritas_kxkt.f:105)."""
    cfg = synth.get_config("cfg1", 0.03)
    cfg.im, cfg.jm = 144, 91
    cfg.kt_list = [44, 4, 5, 1, 2, 3, 44]
    cfg.kx_lists = [[120, 130], [220], [220], [181, 187], [181], [180], [130]]
    files = []
    for i in range(2):
        c = synth.make_file(synth.get_config("cfg1", 0.03), i)
        rng = np.random.default_rng(100 + i)
        row = rng.integers(0, 7, len(c["lat"]))
        c["kt"] = np.asarray(cfg.kt_list, np.int32)[row]
        c["kx"] = np.asarray([120, 220, 220, 181, 181, 180, 130], np.int32)[row]
        surf = np.isin(c["kt"], [1, 2, 3])
        c["lev"] = np.where(surf, rng.choice([0.0, -5.0, 5000.0, np.nan], len(surf)), c["lev"])  # ignored
        files.append(c)
    table, l2d, l3d, *_ = H.setup(cfg)
    assert (l2d, l3d) == (3, 4)
    for nr in (1, 2):
        ref, rf = H.oracle_run(cfg, files, nr)
        out, gf = H.gpu_run(cfg, files, nr, mode)
        assert all(np.array_equal(a, b) for a, b in zip(gf, rf))
        H.compare(out, ref, l2d, l3d, nr, exact=H.is_det(mode), tol=TOL_FAST)


def _one_obs_box(cfg, lat, lon, lev, kx=120, kt=44):
    """Which (i,j,k) box (1-based) a single observation lands in, via the count grid."""
    cols = dict(lat=np.array([lat]), lon=np.array([lon]), lev=np.array([lev]), kx=np.array([kx], np.int32),
                kt=np.array([kt], np.int32), omf=np.array([1.0]), oma=np.array([1.0]), xm=np.array([0.0]))
    out, fl = H.gpu_run(cfg, [cols], 1, cabi.MODE_FAST, norm=False)
    idx = np.argwhere(out["nobs_3d"] == 1.0)
    assert idx.shape[0] == 1
    return tuple(int(v) + 1 for v in idx[0][:3])


def test_kat_box_assignment():
    """Known-answer tests derived from the reference source."""
    cfg = synth.get_config("cfg1", 0.001)
    im, jm = cfg.im, cfg.jm
    p1, p2 = oracle.pint(cfg.plevs)
    dlon = 360.0 / im
    assert _one_obs_box(cfg, 0.0, 180.0, 500.0)[0] == 1                       # i = im+1 -> 1 (m_gritas_grids.f:421)
    assert _one_obs_box(cfg, 0.0, -180.0 - dlon / 2 - 1e-9, 500.0)[0] == im   # i = 0 -> im (:422)
    assert _one_obs_box(cfg, 0.0, 270.0, 500.0)[0] == im                      # clamped, not wrapped (:425)
    assert _one_obs_box(cfg, 0.0, -180.0 + dlon / 2, 500.0)[0] == 2           # nint(0.5) = 1, half away from zero
    assert _one_obs_box(cfg, 0.0, -180.0 + 1.5 * dlon, 500.0)[0] == 3         # nint(1.5) = 2 (banker's would too)
    assert _one_obs_box(cfg, 0.0, -180.0 + 2.5 * dlon, 500.0)[0] == 4         # nint(2.5) = 3 (banker's gives 2)
    assert _one_obs_box(cfg, 90.0, 0.0, 500.0)[1] == jm
    assert _one_obs_box(cfg, -90.0, 0.0, 500.0)[1] == 1
    assert _one_obs_box(cfg, 95.0, 0.0, 500.0)[1] == jm                       # clamp (:426)
    assert _one_obs_box(cfg, -1000.0, 0.0, 500.0)[1] == 1
    k = 5
    assert _one_obs_box(cfg, 0.0, 0.0, p2[k - 1])[2] == k + 1                 # lev == p2(k) -> level k+1 (.gt. p2)
    assert _one_obs_box(cfg, 0.0, 0.0, p1[k - 1])[2] == k                     # lev == p1(k) -> level k (.le. p1)
    assert _one_obs_box(cfg, 0.0, 0.0, 2000.0)[2] == 1
    assert _one_obs_box(cfg, 0.0, 0.0, 1e-300)[2] == cfg.nlevs


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("badlev", [2000.5, 0.0, -3.0, float("nan")])
def test_level_miss_is_fatal(mode, badlev):
    """m_gritas_grids.f:458-459: print the level, gritas_perror -> exit(7).  Only for accepted 3-D obs."""
    cfg, files = small("cfg1", 0.002, 1)
    c = files[0]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fl = np.zeros(len(c["lat"]), np.int32)
    good = np.nonzero((c["kt"] != 7))[0]
    pos = int(good[len(good) // 2])
    c["lev"][pos] = badlev
    # an earlier bad level on a rejected / not-in-table observation must NOT be fatal
    rej = int(np.nonzero(c["kt"] == 7)[0][0])
    c["lev"][rej] = -1.0
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    with pytest.raises(GritasPerror) as ei:
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), len(fl), 1, table, p1, p2, fl,
                       l2d, l3d=l3d)
    assert ei.value.code == 7
    assert (np.isnan(badlev) and np.isnan(ei.value.lev)) or ei.value.lev == badlev
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 1)
    with pytest.raises(oracle.LevelMiss) as eo:
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), table, p1, p2,
                     np.zeros(len(fl), np.int32))
    assert eo.value.n == pos
    # the context stays usable after the error is reported
    c["lev"][pos] = 500.0
    G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), len(fl), 1, table, p1, p2, None,
                   l2d, l3d=l3d)
    G.close()


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("n", [0, 1, 2, 3, 5, 31, 33, 257, 1023])
def test_empty_and_ragged(mode, n):
    cfg = synth.get_config("cfg1", 1.0)
    cfg.im, cfg.jm = 72, 46
    cfg.nobs_per_file = max(n, 1)
    c = synth.make_file(cfg, 7)
    c = {k: v[:n].copy() for k, v in c.items()}
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, rf = H.oracle_run(cfg, [c], 2)
    out, gf = H.gpu_run(cfg, [c], 2, mode)
    assert np.array_equal(gf[0], rf[0])
    H.compare(out, ref, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("mode", H.MODES)
def test_unaligned_columns(mode):
    """Column sections that start on an odd element (8-byte but not 16-byte aligned; 4-byte ints)."""
    cfg, files = small("cfg1", 0.01, 1)
    c = {k: v[1:].copy() for k, v in files[0].items()}
    n = len(c["lat"])
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    back = {k: np.concatenate([v[:1], v]) for k, v in c.items()}
    view = {k: v[1:] for k, v in back.items()}                    # views with a one-element offset
    dfull = np.zeros(2 * n + 1)
    dfull[1:n + 1] = c["omf"]
    dfull[n + 1:] = c["oma"]
    dview = dfull[1:]
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    flfull = np.zeros(n + 1, np.int32)
    G.GrITAS_Accum(view["lat"], view["lon"], view["lev"], view["kx"], view["kt"], dview, n, 2, table, p1, p2,
                   flfull[1:], l2d, l3d=l3d)
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    ref, rf = H.oracle_run(cfg, [c], 2)
    assert np.array_equal(flfull[1:], rf[0])
    H.compare(out, ref, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("mode", H.MODES)
def test_device_resident_inputs(mode):
    cfg, files = small("cfg1", 0.02, 2)
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 2)
    out, _ = H.gpu_run(cfg, files, 2, mode, want_flags=False, device_inputs=True)
    H.compare(out, ref, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)


def test_single_level():
    """nlevs = 1: only plevs(1)-0.1 < lev <= plevs(1) is accepted (m_gritas_grids.f:249-256); Norm
    leaves the 3-D sums raw (returns before the 3-D loop, :607)."""
    cfg = synth.get_config("cfg1", 0.002)
    cfg.plevs = np.array([500.0])
    c = synth.make_file(synth.get_config("cfg1", 0.002), 0)
    c["lev"] = np.full(len(c["lat"]), 499.95)
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, rf = H.oracle_run(cfg, [c], 1)
    out, gf = H.gpu_run(cfg, [c], 1, cabi.MODE_DETERMINISTIC)
    assert np.array_equal(gf[0], rf[0])
    for nm in ("bias_3d", "rtms_3d", "nobs_3d"):
        assert np.array_equal(out[nm], getattr(ref, nm))
    assert np.array_equal(out["xcov_3d"][..., 0], ref.xcov_3d[..., 0])
    assert np.all(out["stdv_3d"] == 0.0)                          # untouched
    c["lev"][3] = oracle.pint(cfg.plevs)[1][0]                    # == p2: rejected (.gt.)
    c["kt"][3], c["kx"][3] = 44, 120
    with pytest.raises(GritasPerror):
        H.gpu_run(cfg, [c], 1, cabi.MODE_FAST)


@pytest.mark.parametrize("nrmlz,ntresh", [(True, None), (True, 0), (True, 1), (False, None)])
def test_norm_variants(nrmlz, ntresh):
    """-nonorm (m stays 1 -> division by m-1 = 0, m_gritas_grids.f:572, 593) and ntresh <= 1."""
    cfg, files = small("cfg1", 0.2, 1)      # dense enough for boxes with n = 0, 1, 2, ...
    cfg2 = synth.get_config("cfg1", 0.2)
    cfg2.im, cfg2.jm = 24, 13
    cfg.im, cfg.jm = 24, 13
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 2, nrmlz=nrmlz, ntresh=ntresh or 0)
    out, _ = H.gpu_run(cfg, files, 2, cabi.MODE_DETERMINISTIC, nrmlz=nrmlz, ntresh=ntresh)
    with np.errstate(invalid="ignore"):
        H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)


def test_norm_ntresh_large():
    """ntresh > 1 (grisas -ntresh): boxes below the threshold are missing.  The reference's 3-D n==1
    branch then reads stale temporaries of an earlier box (m_gritas_grids.f:642-644) -- those boxes
    are excluded from the comparison (documented divergence)."""
    cfg, files = small("cfg1", 0.2, 1)
    cfg.im, cfg.jm = 24, 13
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 1, ntresh=4)
    out, _ = H.gpu_run(cfg, files, 1, cabi.MODE_DETERMINISTIC, ntresh=4)
    n = ref.nobs_3d
    assert np.array_equal(out["nobs_3d"], n)
    for nm in ("bias_3d", "rtms_3d"):
        assert np.array_equal(out[nm], getattr(ref, nm))
    keep = (n != 1.0)[..., None]
    for nm in ("stdv_3d", "xcov_3d"):
        assert np.array_equal(np.where(keep, out[nm], 0.0), np.where(keep, getattr(ref, nm), 0.0))
    assert np.all(out["bias_3d"][(n < 4)[..., None] & np.ones_like(out["bias_3d"], bool)] == oracle.AMISS)


def test_n_equals_one_quirk():
    """Box with n = 1: 2-D stdv = AMISS, 3-D stdv = 0 (m_gritas_grids.f:597-600 vs 642-644)."""
    cfg = synth.get_config("cfg1", 0.001)
    cfg.kt_list, cfg.kx_lists = [44, 1], [[120], [181]]
    cols = dict(lat=np.array([10.0, 10.0]), lon=np.array([20.0, 20.0]), lev=np.array([500.0, 500.0]),
                kx=np.array([120, 181], np.int32), kt=np.array([44, 1], np.int32), omf=np.array([2.0, 3.0]),
                oma=np.array([1.0, 1.0]), xm=np.zeros(2))
    out, _ = H.gpu_run(cfg, [cols], 1, cabi.MODE_DETERMINISTIC)
    i3 = np.argwhere(out["nobs_3d"] == 1.0)[0]
    i2 = np.argwhere(out["nobs_2d"] == 1.0)[0]
    assert out["stdv_3d"][tuple(i3) + (0,)] == 0.0
    assert out["stdv_2d"][tuple(i2) + (0,)] == oracle.AMISS
    assert out["bias_3d"][tuple(i3) + (0,)] == 2.0 and out["bias_2d"][tuple(i2) + (0,)] == 3.0
    assert np.count_nonzero(out["bias_3d"] != oracle.AMISS) == 1       # every empty box is AMISS


def test_three_residuals_raw():
    """nr = 3 (three-corner hat): Accum leaves xcov untouched (nx = 0, m_gritas_grids.f:397-401); Norm
    refuses nr > 2 (:555)."""
    cfg, files = small("cfg1", 0.01, 1)
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 3, norm=False)
    out, _ = H.gpu_run(cfg, files, 3, cabi.MODE_DETERMINISTIC, norm=False)
    for nm in ("bias_3d", "rtms_3d", "nobs_3d"):
        assert np.array_equal(out[nm], getattr(ref, nm))
    assert np.all(out["xcov_3d"][..., 0] == 0.0) and np.all(ref.xcov_3d == 0.0)
    with pytest.raises(GritasPerror):
        H.gpu_run(cfg, files, 3, cabi.MODE_DETERMINISTIC, norm=True)


def test_raw_round_trip_and_reset():
    """get_raw / add_raw (the -nonorm combine workflow, gritas.f:1508-1509) and reset (grisas.f:285-296)."""
    cfg, files = small("cfg1", 0.02, 2)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 2)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_DETERMINISTIC)
    c = files[0]
    G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), len(c["lat"]), 2, table, p1, p2,
                   None, l2d, l3d=l3d)
    raw = G.alloc_grids()
    G.get_raw(None, None, None, None, raw["bias_3d"], raw["rtms_3d"], raw["xcov_3d"], raw["nobs_3d"])
    G.reset()
    zero = G.alloc_grids()
    G.get_raw(None, None, None, None, zero["bias_3d"], zero["rtms_3d"], zero["xcov_3d"], zero["nobs_3d"])
    assert not zero["nobs_3d"].any() and not zero["bias_3d"].any()
    G.add_raw(None, None, None, None, raw["bias_3d"], raw["rtms_3d"], raw["xcov_3d"], raw["nobs_3d"])
    c = files[1]
    G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), len(c["lat"]), 2, table, p1, p2,
                   None, l2d, l3d=l3d)
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("iopt,ioptn", [(1, 1), (1, -1), (3, 3), (3, -3), (101, 101), (101, -101)])
def test_fused_ods_entry(mode, iopt, ioptn):
    """gritas.f:562-587 + 710-739 + GrITAS_Accum_ fused; passive data kept iff ioptn > 0."""
    cfg, files = small("cfg1", 0.03, 2)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    nr = 2 if iopt == 101 else 1
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    for c in files:
        n = len(c["lat"])
        fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=ioptn, t1=-90, t2=120, lona=-100.0, lonb=140.0,
                                 x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        d = np.zeros((n, nr), order="F")
        d[:, 0] = oracle.select_del(1 if iopt == 101 else iopt, c["omf"], c["oma"])
        if nr == 2:
            d[:, 1] = oracle.select_del(3, c["omf"], c["oma"])
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        gfl = np.full(n, -9, np.int32)
        G.accum_ods(c, iopt=iopt, ioptn=ioptn, t1=-90, t2=120, lona=-100.0, lonb=140.0, x_passive=synth.X_PASSIVE,
                    x_pre_bad=synth.X_PRE_BAD, kxkt_table=table, p1=p1, p2=p2, l2d=l2d, l3d=l3d, ob_flag_out=gfl)
        assert np.array_equal(gfl, fl)
    oracle.norm(g, cfg.nlevs, True, 0)
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    H.compare(out, g, l2d, l3d, nr, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("scale_res", [1, 2, 3])
def test_fused_ods_scaling(scale_res):
    """gritas.f:564-572: del/sqrt(xvec), del/obs, del/(obs-omf)."""
    cfg, files = small("cfg1", 0.02, 1)
    c = files[0]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    for iopt in (1, 3):
        g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 1)
        fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=iopt, x_passive=synth.X_PASSIVE,
                                 x_pre_bad=synth.X_PRE_BAD)
        d = oracle.select_del(iopt, c["omf"], c["oma"], c["obs"], c["xvec"], scale_res)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        oracle.norm(g, cfg.nlevs, True, 0)
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_DETERMINISTIC)
        G.accum_ods(c, iopt=iopt, scale_res=scale_res, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD,
                    kxkt_table=table, p1=p1, p2=p2, l2d=l2d, l3d=l3d)
        out = G.alloc_grids()
        G.norm_into(out)
        G.close()
        H.compare(out, g, l2d, l3d, 1, exact=True, tol=0.0)


@pytest.mark.parametrize("name,scale", [("cfg3_amsua", 0.002), ("cfg3_iasi", 0.004), ("cfg4", 0.03), ("cfg5", 0.05)])
@pytest.mark.parametrize("mode", H.MODES)
def test_other_configs_small(name, scale, mode):
    """BASELINE.json configs 3-5 at sizes the oracle finishes in seconds (grid shrunk for radiances:
    the oracle's host grids at 0.25 deg x 616 channels would need 20 GB)."""
    cfg = synth.get_config(name, scale)
    if cfg.kind == "radiance":
        cfg.im, cfg.jm = 144, 73
    if name == "cfg4":
        cfg.im, cfg.jm = 36, 19
    files = [synth.make_file(cfg, i) for i in range(2)]
    table, l2d, l3d, *_ = H.setup(cfg)
    nr = cfg.nr
    ref, rf = H.oracle_run(cfg, files, nr)
    out, gf = H.gpu_run(cfg, files, nr, mode)
    assert all(np.array_equal(a, b) for a, b in zip(gf, rf))
    H.compare(out, ref, l2d, l3d, nr, exact=H.is_det(mode), tol=TOL_FAST)


def test_deterministic_is_reproducible_on_clustered_data():
    """config 5: two runs of the deterministic path must be bit-identical; the atomic path need not be."""
    cfg, files = small("cfg5", 0.25, 2)
    a, _ = H.gpu_run(cfg, files, 1, cabi.MODE_DETERMINISTIC, want_flags=False)
    b, _ = H.gpu_run(cfg, files, 1, cabi.MODE_DETERMINISTIC, want_flags=False)
    for nm in H.NAMES:
        assert np.array_equal(a[nm], b[nm], equal_nan=True)
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 1)
    H.compare(a, ref, l2d, l3d, 1, exact=True, tol=0.0)
    f, _ = H.gpu_run(cfg, files, 1, cabi.MODE_FAST, want_flags=False)
    H.compare(f, ref, l2d, l3d, 1, exact=False, tol=TOL_FAST)


def test_full_size_file_properties():
    """One full-size cfg2 file (1.21 M obs, omf+oma): size-independent properties + full oracle check."""
    cfg = synth.get_config("cfg2")
    c = synth.make_file(cfg, 3)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=101, x_passive=synth.X_PASSIVE,
                             x_pre_bad=synth.X_PRE_BAD)
    results = {}
    for mode in (cabi.MODE_FAST, cabi.MODE_DETERMINISTIC):
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
        gfl = np.zeros(len(fl), np.int32)
        G.accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD, kxkt_table=table, p1=p1, p2=p2,
                    l2d=l2d, l3d=l3d, ob_flag_out=gfl)
        raw = G.alloc_grids()
        G.get_raw(None, None, None, None, raw["bias_3d"], raw["rtms_3d"], raw["xcov_3d"], raw["nobs_3d"])
        G.close()
        accepted = gfl == 0
        assert raw["nobs_3d"].sum() == accepted.sum()                            # checksum of counts
        assert np.isclose(raw["bias_3d"][..., 0].sum(), c["omf"][accepted].sum(), rtol=1e-9, atol=1e-6)
        assert np.isclose(raw["rtms_3d"][..., 1].sum(), (c["oma"][accepted] ** 2).sum(), rtol=1e-12)
        results[mode] = raw
    assert np.array_equal(results[0]["nobs_3d"], results[1]["nobs_3d"])
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 2)
    d = np.zeros((len(fl), 2), order="F")
    d[:, 0], d[:, 1] = c["omf"], c["oma"]
    oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
    for nm in ("bias_3d", "rtms_3d", "nobs_3d"):
        assert np.array_equal(results[1][nm], getattr(g, nm))
    assert np.array_equal(results[1]["xcov_3d"][..., 0], g.xcov_3d[..., 0])


@pytest.mark.parametrize("name", ["cfg1", "cfg5"])
def test_staging_overflow_falls_back_to_direct_accumulate(name, monkeypatch):
    """Tile lists with a tiny capacity: most entries overflow and take the direct fp64-RED path inside
    K1 while the rest goes through the tiled accumulate -- the sum of both must still be right."""
    monkeypatch.setenv("GRITAS_B200_STAGE_SLOTS", str(64 * 1024))
    monkeypatch.setenv("GRITAS_B200_TILE_SHIFT", "10")
    cfg, files = small(name, 0.1, 3, res=(24, 13))
    table, l2d, l3d, *_ = H.setup(cfg)
    for nr in (1, 2):
        ref, rf = H.oracle_run(cfg, files, nr)
        out, gf = H.gpu_run(cfg, files, nr, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED)
        assert all(np.array_equal(a, b) for a, b in zip(gf, rf))
        H.compare(out, ref, l2d, l3d, nr, exact=False, tol=TOL_FAST)


def test_staged_flush_between_files(monkeypatch):
    """Capacity reached after every file -> the lists are flushed between Accum calls."""
    monkeypatch.setenv("GRITAS_B200_STAGE_SLOTS", str(256 * 1024))
    cfg, files = small("cfg2", 0.1, 4)
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 2)
    out, _ = H.gpu_run(cfg, files, 2, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, want_flags=False)
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=TOL_FAST)


@pytest.mark.parametrize("name,expect", [("cfg2", 1), ("cfg3_amsua", 0)])
def test_fast_mode_picks_the_path_from_the_data(name, expect):
    """Plain fast mode: the first call with >= 64 Ki observations runs the direct kernel with a locality probe;
    scattered conventional data then switches to the tiled accumulate, radiance footprints (consecutive channels
    -> consecutive records) stay on the direct path.  Results must not depend on the choice."""
    cfg = synth.get_config(name, 0.1 if name == "cfg2" else 0.006)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(3)]
    assert len(files[0]["lat"]) >= 65536
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    nr = cfg.nr
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_FAST)
    for c in files:
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, nr), len(c["lat"]), nr, table, p1, p2,
                       None, l2d, l3d=l3d)
    assert cabi.load().gritas_b200_tiled(G.handle) == expect
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    ref, _ = H.oracle_run(cfg, files, nr)
    H.compare(out, ref, l2d, l3d, nr, exact=False, tol=TOL_FAST)


@pytest.mark.parametrize("mode", [cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, cabi.MODE_DETERMINISTIC])
@pytest.mark.parametrize("iopt,options", [(5, 0), (7, 0), (9, 0), (11, 0), (11, cabi.OPT_TCH), (13, 0), (13, cabi.OPT_TCH),
                                          (15, 0), (15, cabi.OPT_TCH), (15, cabi.OPT_TCH | cabi.OPT_SELF), (-17, 0), (19, 0),
                                          (21, 0), (23, 0), (-25, 0), (29, 0), (31, 0), (33, 0), (16, 0), (32, 0)])
def test_fused_entry_all_residual_options(mode, iopt, options):
    """gritas_b200_accum_odsx: the residual expressions of gritas.f:588-704 evaluated on the device must be
    bit-identical to the oracle's host-side del (deterministic mode) for every single-ODS option."""
    cfg, files = small("cfg1", 0.02, 2)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    tch, self_ = bool(options & cabi.OPT_TCH), bool(options & cabi.OPT_SELF)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    g = None
    for c in files:
        c = dict(c)
        c["xvec"] = np.where(np.arange(len(c["lat"])) % 97 == 0, -1.0, c["xvec"])     # non-positive sigo now and then
        d = oracle.select_del_all(iopt, c["omf"], c["oma"], c["obs"], c["xvec"], c["xm"], tch=tch, self_=self_)
        nr = d.shape[1]
        if g is None:
            g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
        n = len(c["lat"])
        if abs(iopt) & 1:          # the QC filter only runs for odd options (gritas.f:710)
            fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=iopt, x_passive=synth.X_PASSIVE,
                                     x_pre_bad=synth.X_PRE_BAD)
        else:
            fl = np.zeros(n, np.int32)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        gfl = np.full(n, -9, np.int32)
        G.accum_odsx(c, iopt, options=options, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD, kxkt_table=table,
                     p1=p1, p2=p2, l2d=l2d, l3d=l3d, ob_flag_out=gfl)
        assert np.array_equal(gfl, fl)
    out = G.alloc_grids()
    raw = nr == 3
    if raw:
        G.get_raw(None, None, None, None, out["bias_3d"], out["rtms_3d"], out["xcov_3d"], out["nobs_3d"])
    else:
        oracle.norm(g, cfg.nlevs, True, 0)
        G.norm_into(out)
    G.close()
    with np.errstate(invalid="ignore"):
        if raw and H.is_det(mode):
            for nm in ("bias_3d", "rtms_3d", "nobs_3d"):
                assert np.array_equal(out[nm], getattr(g, nm))
        elif raw:
            assert np.array_equal(out["nobs_3d"], g.nobs_3d)
            assert np.allclose(out["bias_3d"], g.bias_3d, rtol=1e-9, atol=1e-9)
        else:
            H.compare(out, g, l2d, l3d, nr, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("mode", [cabi.MODE_FAST, cabi.MODE_DETERMINISTIC])
def test_minmax_matches_oracle_exactly(mode):
    """GrITAS_MinMax_ (m_gritas_grids.f:1019-1167, -minmax): min / max / count per box, surface and upper air;
    min and max are order-free, so every mode must be bit-exact.  NaN residuals count but never win."""
    cfg = synth.get_config("cfg1", 0.05)
    cfg.im, cfg.jm = 72, 46
    cfg.kt_list = [44, 4, 1, 2]
    cfg.kx_lists = [[120, 130], [220], [181], [181, 187]]
    files = []
    for i in range(2):
        c = synth.make_file(synth.get_config("cfg1", 0.05), i)
        rng = np.random.default_rng(300 + i)
        pairs = [(44, 120), (44, 130), (4, 220), (1, 181), (2, 181), (2, 187), (7, 120)]
        pick = rng.integers(0, len(pairs), len(c["lat"]))
        c["kt"] = np.asarray([p[0] for p in pairs], np.int32)[pick]
        c["kx"] = np.asarray([p[1] for p in pairs], np.int32)[pick]
        c["omf"][::53] = np.nan
        c["omf"][1::211] = 0.0
        c["omf"][2::211] = -0.0
        files.append(c)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 1)
    oracle.minmax_init(g)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode | cabi.FLAG_MINMAX)
    for c in files:
        fl = H.filter_in(c)
        gfl = fl.copy()
        oracle.minmax(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], c["omf"], table, p1, p2, fl)
        G.GrITAS_MinMax(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], np.asfortranarray(c["omf"].reshape(-1, 1)),
                        len(fl), 1, table, p1, p2, gfl, l2d, l3d=l3d)
        assert np.array_equal(gfl, fl)
    out = G.alloc_grids()
    G.get_minmax(out["bias_2d"], out["stdv_2d"], out["nobs_2d"], out["bias_3d"], out["stdv_3d"], out["nobs_3d"])
    with pytest.raises(RuntimeError):
        G.norm_into(G.alloc_grids())           # the driver never calls Norm with -minmax (gritas.f:884)
    G.close()
    for nm in ("bias_2d", "stdv_2d", "nobs_2d", "bias_3d", "stdv_3d", "nobs_3d"):
        assert np.array_equal(out[nm], getattr(g, nm)), nm
    assert (out["bias_3d"] == oracle.AMISS).any() and (out["stdv_3d"] == -oracle.AMISS).any()
    with pytest.raises(Exception):
        Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.FLAG_MINMAX)._ensure(2, table, p1, p2, l2d, l3d)


@pytest.mark.parametrize("mode", H.MODES)
def test_mask_gather_matches_maskout(mode):
    """gritas_maskout (m_gritas_masks.F90:137-186) fused into the accumulate kernels."""
    cfg, files = small("cfg2", 0.05, 2)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    rng = np.random.default_rng(5)
    maskfld = np.asfortranarray(rng.choice([0.0, 0.5, 0.98999, 0.99, 1.0], size=(cfg.im, cfg.jm)))
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 2)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    G._ensure(2, table, p1, p2, l2d, l3d)
    G.set_mask(maskfld)
    for c in files:
        fl = H.filter_in(c, 101)
        nex = oracle.maskout(cfg.im, cfg.jm, c["lat"], c["lon"], maskfld, fl)       # gritas.f:745
        assert 0 < nex < len(fl)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), table, p1, p2, fl)
        gfl = np.full(len(fl), -9, np.int32)
        G.accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD, ob_flag_out=gfl)
        assert np.array_equal(gfl, fl)
    oracle.norm(g, cfg.nlevs, True, 0)
    out = G.alloc_grids()
    G.norm_into(out)
    G.set_mask(None)
    G.close()
    H.compare(out, g, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("nr", [1, 2])
def test_fused_norm_consumes_the_tile_lists_without_changing_the_state(nr, monkeypatch):
    """Tiled path, one rank: GrITAS_Norm runs K2f (tile accumulate fused with finalize) on the parked lists.  It must
    (a) agree with the oracle, (b) be repeatable -- Norm does not change the accumulation state, (c) let the
    accumulation continue afterwards, (d) agree with the unfused path (flush into the records, then K3), including when
    part of the data was already flushed into the records, and (e) honour a reset."""
    cfg, files = small("cfg2", 0.05, 4)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    mode = cabi.MODE_FAST | cabi.FLAG_FORCE_TILED
    L = cabi.load()

    def accum(G, fs):
        for c in fs:
            G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, nr), len(c["lat"]), nr, table, p1, p2,
                           None, l2d, l3d=l3d)

    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    accum(G, files[:2])
    assert L.gritas_b200_fused_norm(G.handle) == 1
    a, b = G.alloc_grids(), G.alloc_grids()
    G.norm_into(a)
    assert L.gritas_b200_fused_norm(G.handle) == 1            # the lists are still there
    G.norm_into(b)
    ref2, _ = H.oracle_run(cfg, files[:2], nr)
    H.compare(a, ref2, l2d, l3d, nr, exact=False, tol=TOL_FAST)
    for nm in H.NAMES:
        if nm.startswith("nobs"):
            assert np.array_equal(a[nm], b[nm]), nm
        else:                                                  # the order of the additions inside a box is not fixed
            assert np.allclose(a[nm], b[nm], rtol=1e-9, atol=1e-9, equal_nan=True), nm
    G.flush()                                                  # half of the month now sits in the records
    assert L.gritas_b200_fused_norm(G.handle) == 0
    accum(G, files[2:])
    assert L.gritas_b200_fused_norm(G.handle) == 1
    c4 = G.alloc_grids()
    G.norm_into(c4)                                            # K2f adds the records to the tile sums
    ref4, _ = H.oracle_run(cfg, files, nr)
    H.compare(c4, ref4, l2d, l3d, nr, exact=False, tol=TOL_FAST)
    raw = G.alloc_grids()
    G.get_raw(raw["bias_2d"] if l2d else None, raw["rtms_2d"] if l2d else None, raw["xcov_2d"] if l2d else None,
              raw["nobs_2d"] if l2d else None, raw["bias_3d"] if l3d else None, raw["rtms_3d"] if l3d else None,
              raw["xcov_3d"] if l3d else None, raw["nobs_3d"] if l3d else None)
    refraw, _ = H.oracle_run(cfg, files, nr, norm=False)
    H.compare(raw, refraw, l2d, l3d, nr, exact=False, tol=TOL_FAST, raw=True)
    G.reset()
    accum(G, files[:1])
    r1 = G.alloc_grids()
    G.norm_into(r1)
    ref1, _ = H.oracle_run(cfg, files[:1], nr)
    H.compare(r1, ref1, l2d, l3d, nr, exact=False, tol=TOL_FAST)
    G.close()
    # unfused path on the same data
    monkeypatch.setenv("GRITAS_B200_FUSED_NORM", "0")
    out, _ = H.gpu_run(cfg, files, nr, mode, want_flags=False)
    assert np.array_equal(out["nobs_3d"], c4["nobs_3d"])
    H.compare(out, ref4, l2d, l3d, nr, exact=False, tol=TOL_FAST)


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("nrmlz,ntresh,km_extra", [(True, None, 0), (False, None, 0), (True, 3, 2)])
def test_3ch_norm_matches_the_oracle(mode, nrmlz, ntresh, km_extra):
    """-tch: three residual series (omf, oma, xm) accumulated with nr = 3 and finalized by GrITAS_3CH_Norm_
    (m_gritas_grids.f:1178-1341; gritas.f:813-818): variances in the stdv slot, 3CH estimates in xcov.  Bit-exact in
    deterministic mode; km > nlevs leaves the upper levels untouched."""
    cfg, files = small("cfg1", 0.02, 2)
    km = cfg.km + km_extra
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, km, l2d, l3d, 3)
    G = Grids(cfg.im, cfg.jm, km, cfg.plevs, mode=mode)
    for c in files:
        d = H.del_of(c, 3)
        fl = np.zeros(len(c["lat"]), np.int32)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, len(c["lat"]), 3, table, p1, p2, None, l2d, l3d=l3d)
    oracle.norm_3ch(g, cfg.nlevs, nrmlz, 0 if ntresh is None else ntresh)
    out = G.alloc_grids()
    with pytest.raises(Exception):
        G.norm_into(out)                                   # GrITAS_Norm_ rejects nr = 3 (m_gritas_grids.f:555)
    G.GrITAS_3CH_Norm(3, nrmlz, l2d, out["bias_2d"] if l2d else None, out["rtms_2d"] if l2d else None,
                      out["stdv_2d"] if l2d else None, out["xcov_2d"] if l2d else None, out["nobs_2d"] if l2d else None,
                      l3d, out["bias_3d"], out["rtms_3d"], out["stdv_3d"], out["xcov_3d"], out["nobs_3d"], ntresh)
    G.close()
    single = g.nobs_3d == 1.0
    for tag, on in (("2d", l2d), ("3d", l3d)):
        if not on:
            continue
        assert np.array_equal(out["nobs_" + tag], getattr(g, "nobs_" + tag))
        for nm in ("bias", "rtms", "stdv", "xcov"):
            a, b = out[f"{nm}_{tag}"], getattr(g, f"{nm}_{tag}")
            if ntresh and tag == "3d" and nm in ("stdv", "xcov"):
                # ntresh > 1 and a box with one observation: the reference reads a stale temporary (documented divergence)
                a, b = a[~single], b[~single]
            if H.is_det(mode):
                assert np.array_equal(a, b, equal_nan=True), f"{nm}_{tag}"
            else:
                big = np.abs(b) > 1e14                      # AMISS and the 3CH combinations of AMISS
                assert np.array_equal(a[big], b[big]), f"{nm}_{tag} missing pattern"
                assert np.allclose(a[~big], b[~big], rtol=1e-6, atol=1e-9, equal_nan=True), f"{nm}_{tag}"


@pytest.mark.parametrize("mode", H.MODES)
def test_per_time_mode_reset_accum_norm_each_synoptic_time(mode):
    """grisas.f:279-689: the grids are zeroed, accumulated and normalised (with ntresh) for every synoptic time
    separately; only the means, the std-dev slot and the counts are written (grisas.f:663-675), so the other output
    arrays are not requested.  One handle serves all times (gritas_b200_reset)."""
    cfg, files = small("cfg1", 0.05, 3)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    for c in files:
        ref, _ = H.oracle_run(cfg, [c], 1, ntresh=2)
        if G.handle is not None:
            G.reset()                                          # grisas.f:285-296
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), len(c["lat"]), 1, table, p1, p2,
                       None, l2d, l3d=l3d)
        out = G.alloc_grids()
        G.GrITAS_Norm(1, True, l2d, out["bias_2d"] if l2d else None, None, out["stdv_2d"] if l2d else None, None,
                      out["nobs_2d"] if l2d else None, l3d, out["bias_3d"], None, out["stdv_3d"], None, out["nobs_3d"], ntresh=2)
        single = ref.nobs_3d == 1.0                            # ntresh > 1, n = 1: stale temporaries in the reference
        for tag, on in (("2d", l2d), ("3d", l3d)):
            if not on:
                continue
            assert np.array_equal(out["nobs_" + tag], getattr(ref, "nobs_" + tag))
            for nm in ("bias", "stdv"):
                a, b = out[f"{nm}_{tag}"], getattr(ref, f"{nm}_{tag}")
                if tag == "3d" and nm == "stdv":
                    a, b = a[~single], b[~single]
                assert np.array_equal(a == oracle.AMISS, b == oracle.AMISS)
                if H.is_det(mode):
                    assert np.array_equal(a, b)
                else:
                    assert np.allclose(a, b, rtol=1e-6, atol=1e-9)
            assert not out[f"rtms_{tag}"].any() and not out[f"xcov_{tag}"].any()      # not requested, not written
    G.close()


@pytest.mark.parametrize("name,device_inputs", [("cfg1", False), ("cfg5", False), ("cfg3_amsua", True)])
def test_presort_is_the_reference_permutation(name, device_inputs):
    """gritas_b200_presort = IndexSet + the eight stable IndexSorts of gritas.f:499-508: the permutation must equal the
    oracle's merge-sort restatement index for index (integer-exact), with duplicates (clustered sites: identical lat /
    lon), signed zeros and negative keys in the columns, before any handle exists and through an existing handle."""
    cfg = synth.get_config(name, 0.05 if name != "cfg3_amsua" else 0.004)
    c = dict(synth.make_file(cfg, 3))
    n = len(c["lat"])
    lat, lon = c["lat"].copy(), c["lon"].copy()
    lat[::17] = 0.0
    lat[5::34] = -0.0                                   # equal to +0.0 in the reference's `<` merge
    lon[::13] = -lon[::13]
    time = c["time"].copy()
    time[::7] = -time[::7]                              # negative int keys
    cols = dict(qcexcl=c["qcexcl"], lat=lat, lon=lon, lev=c["lev"], time=time, kt=c["kt"], ks=c["ks"], kx=c["kx"])
    ref = oracle.presort(**cols)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs)        # no handle yet: the library's scratch context
    args = cols
    keep = None
    if device_inputs:
        import torch
        keep = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in cols.items()}
        args = keep
    got0 = G.presort(**args)
    assert np.array_equal(got0, ref)
    got1 = G.presort(**args, index_base=1)              # Fortran's `is`
    assert np.array_equal(got1, ref + 1)
    # sorted columns are in (kx, ks, kt, time, lev, lon, lat, qcexcl) order
    k = np.stack([cols[nm][got0] for nm in ("kx", "ks", "kt", "time", "lev", "lon", "lat", "qcexcl")], axis=1).astype(np.float64)
    d = np.diff(k, axis=0)
    first = np.argmax(d != 0, axis=1)
    assert np.all((d[np.arange(n - 1), first] > 0) | ~(d != 0).any(axis=1))
    G.close()


@pytest.mark.parametrize("case", ["plain", "overflow", "flushed", "records_only", "direct"])
def test_multi_rank_norm_pipeline_on_one_rank(case, monkeypatch):
    """The multi-rank Norm pipeline without the collective (GRITAS_B200_FORCE_PIPELINE=1): tiled path = K2x (tile sums
    exported as compact planes, records folded in where a tile has any) -> K3x (finalize from the planes); other paths =
    tile flush -> K3.  'overflow': tiny lists, most entries take the direct accumulate (dirty tiles); 'flushed': half of
    the data already sits in the records; 'records_only': the lists exist but the data took the direct kernel (files
    too small for the locality probe) -- the planes are exported from the records; 'direct': no tile lists at all."""
    monkeypatch.setenv("GRITAS_B200_FORCE_PIPELINE", "1")
    if case == "overflow":
        monkeypatch.setenv("GRITAS_B200_STAGE_SLOTS", str(64 * 1024))
    cfg, files = small("cfg2", 0.05, 4)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    mode = cabi.MODE_FAST | {"direct": cabi.FLAG_NO_STAGING, "records_only": 0}.get(case, cabi.FLAG_FORCE_TILED)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    for fi, c in enumerate(files):
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), len(c["lat"]), 2, table, p1, p2,
                       None, l2d, l3d=l3d)
        if case == "flushed" and fi == 1:
            G.flush()
    out = G.alloc_grids()
    G.norm_reduce_into(out, root=0)
    ref, _ = H.oracle_run(cfg, files, 2)
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=TOL_FAST)
    if case != "direct":                                   # the plane-based pipeline does not change the state: Norm still works
        again = G.alloc_grids()
        G.norm_into(again)
        H.compare(again, ref, l2d, l3d, 2, exact=False, tol=TOL_FAST)
    G.close()


def test_no_writeback_flag_leaves_ob_flag_untouched():
    """GRITAS_B200_FLAG_NO_WRITEBACK: ob_flag is an input only (what a driver without -lb needs, gritas.f:791-796);
    the grids are the same as with the write-back."""
    cfg, files = small("cfg1", 0.05, 2)
    fin = [flags_of(cfg, c) for c in files]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref, rflags = H.oracle_run(cfg, files, 1, fin)
    assert any(not np.array_equal(a, b) for a, b in zip(fin, rflags))       # the oracle does set flags (not-in-table obs)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_DETERMINISTIC | cabi.FLAG_NO_WRITEBACK | cabi.FLAG_ASYNC)
    for c, f0 in zip(files, fin):
        fl = f0.copy()
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), len(fl), 1, table, p1, p2, fl, l2d, l3d=l3d)
        assert np.array_equal(fl, f0)
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    H.compare(out, ref, l2d, l3d, 1, exact=True, tol=0.0)


def test_pageable_columns_through_the_copy_pool_full_size():
    """One full-size file (1.21 M obs, ~10 MB per column) from pageable arrays: the staging copies run chunk by chunk on
    the library's host threads, the grids and the ob_flag write-back come back through the pinned bounce ring."""
    cfg = synth.get_config("cfg2", 1.0)
    cfg.im, cfg.jm = 72, 46
    c = synth.make_file(cfg, 3)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fin = H.filter_in(c, 101)
    ref, rflags = H.oracle_run(cfg, [c], 2, [fin])
    for mode in (cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, cabi.MODE_DETERMINISTIC):
        out, gflags = H.gpu_run(cfg, [c], 2, mode, [fin])
        assert np.array_equal(gflags[0], rflags[0])
        H.compare(out, ref, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)


@pytest.mark.parametrize("flags", [0, cabi.FLAG_ASYNC])
def test_early_ob_flag_write_back_pinned_and_pageable(flags):
    """Host callers get ob_flag back while the fp64 columns are still being uploaded (k_early_flags, from kx / kt / the
    incoming flags alone): same flags and statistics as the oracle for pageable and for pinned columns, synchronous and
    ASYNC calls, several files in a row (slot reuse), incoming flags partly set."""
    import torch
    cfg = synth.get_config("cfg2", 0.25)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(5)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fins = [H.filter_in(c, 101) for c in files]
    ref, rflags = H.oracle_run(cfg, files, 2, fins)
    for pinned in (False, True):
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_DETERMINISTIC | flags)
        got = []
        keep = []
        for c, fin in zip(files, fins):
            n = len(c["lat"])
            d = H.del_of(c, 2)
            cols = [c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], np.ascontiguousarray(d.ravel(order="F")), fin.copy()]
            if pinned:
                cols = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in cols]
            keep.append(cols)
            G.GrITAS_Accum(*cols[:6], n, 2, table, p1, p2, cols[6], l2d, l3d=l3d)
            got.append(cols[6])
        out = G.alloc_grids()
        G.norm_into(out)
        G.close()
        for g, r in zip(got, rflags):
            assert np.array_equal(g, r)
        H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("device_inputs", [False, True])
def test_fused_presort_gather_accum_matches_the_sorted_reference_loop(mode, device_inputs):
    """gritas_b200_accum_odsx_sorted = gritas.f:499-523 (8-key pre-sort + the column gathers) + 562-739 + GrITAS_Accum_, all on
    the device from UNSORTED columns.  Oracle: sort on the host, gather, filter, Accum in sorted order; in deterministic mode
    the GPU sums are bit-identical to that sorted sequential loop; ob_flag and `is` come back in the reference's form."""
    cfg, files = small("cfg5", 0.05, 2)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 2)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    G._ensure(2, table, p1, p2, l2d, l3d)
    keep = []
    for c in files:
        is0 = oracle.presort(c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], c["ks"], c["kx"])
        s = {k: np.ascontiguousarray(v[is0]) for k, v in c.items() if isinstance(v, np.ndarray) and v.shape == c["lat"].shape}
        fl = H.filter_in(s, 101)
        oracle.accum(ref, cfg.nlevs, s["lat"], s["lon"], s["lev"], s["kx"], s["kt"], H.del_of(s, 2), table, p1, p2, fl)
        cols = c
        if device_inputs:
            import torch
            cols = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in c.items() if isinstance(v, np.ndarray)}
            keep.append(cols)
        fo = np.full(len(fl), -5, np.int32)
        iso = np.zeros(len(fl), np.int32)
        G.accum_odsx_sorted(cols, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD, ob_flag_out=fo, is_out=iso)
        assert np.array_equal(iso, is0 + 1)
        assert np.array_equal(fo, fl)
    oracle.norm(ref, cfg.nlevs, True, 0)
    out = G.alloc_grids()
    G.norm_into(out)
    G.close()
    H.compare(out, ref, l2d, l3d, 2, exact=H.is_det(mode), tol=TOL_FAST)
