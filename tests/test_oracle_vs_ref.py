"""The hand-written CPU oracle (oracle/gritas_oracle.c) against the reference's OWN Fortran routines, translated to C
statement by statement (oracle/f2c_subset.py reads the sources where they lie under /root/reference; oracle/ref.py loads
the compiled result).  Everything must agree BIT FOR BIT: level intervals, ob_flag, raw sums, statistics, min / max, the
three-cornered-hat Norm, the channel / level cross-covariances, the (kx,kt) table, every residual option, the QC filter,
the mask gather, the pre-sort permutation, the scorecard means and region rows, the mesh sizes, and the statement at which
gritas_perror is called.

This is what pins the oracle: m_gritas_grids.f GrITAS_Pint_ (:217-296), GrITAS_Accum_ (:307-478), GrITAS_Norm_ (:490-656),
GrITAS_XCov_Accum_ (:667-848), GrITAS_XCov_Norm_ (:859-1008), GrITAS_MinMax_ (:1019-1167), GrITAS_3CH_Norm_ (:1178-1341),
scores2_ (:1671-1795), ave_nomiss (:1797-1826), GrITAS_XCov_Prs_Accum_ (:1862-2059), init_ (:128-129, 136-138, 163-175); gritas_kxkt.f GrITAS_KxKt
(:10-119); m_gritas_masks.F90 maskout_ (:152-153, 164-183); gritas.f :500-508, :562-704, :712-735, :738.  Skipped where
neither the reference sources nor a built oracle/_ref/libgritas_ref.so exist."""
import numpy as np
import pytest

from gritas_b200 import synth
from oracle import oracle, ref
from tests import helpers as H

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref is not built and /root/reference is not here")

KT_LIST = [44, 4, 1, 2, 5]
KX_LISTS = [[10, 11], [12], [10], [13, 14], [12]]


def random_case(seed, nr, n=4000, nlev=None):
    rng = np.random.default_rng(seed)
    im, jm = int(rng.integers(4, 40)), int(rng.integers(3, 30))
    nlev = nlev or int(rng.integers(2, 12))
    plevs = np.sort(rng.uniform(1.0, 1050.0, nlev))[::-1].copy()
    table, _, _, l2d, l3d = oracle.kxkt(60, 40, KT_LIST, KX_LISTS)
    lat = rng.uniform(-95, 95, n)
    lon = rng.uniform(-200, 380, n)            # outside [-180, 180): the reference clamps, it does not wrap 0..360
    kx = rng.integers(8, 17, n).astype(np.int32)
    kt = rng.choice(np.array([44, 4, 1, 2, 5, 7, 3, 61], np.int32), n)
    kt = np.clip(kt, 1, 40).astype(np.int32)   # the reference indexes kxkt_table(kx, kt) unchecked
    d = np.asfortranarray(rng.normal(0.3, 2.0, (n, nr)))
    fl0 = (rng.random(n) < 0.2).astype(np.int32)
    # grid-aligned values exercise NINT ties and the +-180 wrap
    lon[:200] = -180.0 + (rng.integers(0, 2 * im, 200) + 0.5) * (360.0 / im)
    lat[:200] = -90.0 + (rng.integers(0, 2 * jm, 200) * 0.5) * (180.0 / (jm - 1))
    lon[200:210] = [-180.0, 180.0, 179.999999, -179.999999, 0.0, 360.0, -360.0, 180.0 - 180.0 / im, -180.0 + 180.0 / im, 1e-300]
    return rng, im, jm, nlev, plevs, table, l2d, l3d, lat, lon, kx, kt, d, fl0


def both(im, jm, km, nlev, plevs, table, l2d, l3d, nr):
    ref.init(im, jm, km, nlev, plevs, table.shape[0], table.shape[1])
    return oracle.Grids(im, jm, km, l2d, l3d, nr), oracle.Grids(im, jm, km, l2d, l3d, nr)


def same(go, gr, l2d, l3d, what):
    for nm in go.names():
        if (nm.endswith("2d") and l2d == 0) or (nm.endswith("3d") and l3d == 0):
            continue
        a, b = getattr(go, nm), getattr(gr, nm)
        assert np.array_equal(a, b, equal_nan=True), f"{what}: {nm} differs at {np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))[:3]}"


@pytest.mark.parametrize("name", H.GOLDEN)
def test_golden_inputs_through_the_translated_reference(name):
    """The committed golden fixtures ARE the outputs of the reference's routines: the translation reproduces every array
    of every fixture bit for bit (tests/test_gpu_golden.py then holds the CUDA path to the same arrays)."""
    cfg, files, flags, norm, raw, l2d, l3d, p1, p2 = H.load_golden(name)
    table, _, _ = synth.kxkt_table(cfg)
    _, g = both(cfg.im, cfg.jm, cfg.km, cfg.nlevs, cfg.plevs, table, l2d, l3d, cfg.nr)
    r1, r2 = ref.pint(cfg.nlevs)
    assert np.array_equal(r1, p1) and np.array_equal(r2, p2)
    for c, fl_gold in zip(files, flags):
        fl = H.filter_in(c)
        ref.accum(g, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, cfg.nr), table, r1, r2, fl)
        assert np.array_equal(fl, fl_gold)
    for nm, a in raw.items():
        assert np.array_equal(getattr(g, nm), a), nm
    ref.norm(g, True, None)
    for nm, a in norm.items():
        if (nm.endswith("2d") and not l2d) or (nm.endswith("3d") and not l3d):
            continue
        assert np.array_equal(getattr(g, nm), a), nm


def test_mesh_sizes():
    for im, jm in ((360, 181), (1440, 721), (72, 46), (7, 3)):
        ref.init(im, jm, 1, 1, np.array([500.0]), 4, 4)
        dlon, dlat = ref.mesh()
        assert dlon == 360.0 / im and dlat == 180.0 / (jm - 1)          # m_gritas_grids.f:128-129


@pytest.mark.parametrize("nlev", [1, 2, 3, 26, 72])
def test_pint(nlev):
    rng = np.random.default_rng(nlev)
    plevs = np.sort(rng.uniform(0.5, 1050.0, nlev))[::-1].copy()
    ref.init(8, 5, nlev, nlev, plevs, 4, 4)
    r1, r2 = ref.pint(nlev)
    o1, o2 = oracle.pint(plevs)
    assert np.array_equal(r1, o1) and np.array_equal(r2, o2)


def test_pint_errors_are_the_same_statements():
    ref.init(8, 5, 3, 3, np.array([500.0, 700.0, 100.0]), 4, 4)          # not decreasing
    with pytest.raises(ref.RefError) as e:
        ref.pint(3)
    assert 260 <= e.value.line <= 270
    with pytest.raises(Exception):
        oracle.pint(np.array([500.0, 700.0, 100.0]))
    ref.init(8, 5, 3, 3, np.array([500.0, 100.0, 0.0]), 4, 4)            # 0 hPa
    with pytest.raises(ref.RefError):
        ref.pint(3)
    with pytest.raises(Exception):
        oracle.pint(np.array([500.0, 100.0, 0.0]))


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("nr", [1, 2, 3])
def test_accum_and_norm(seed, nr):
    rng, im, jm, nlev, plevs, table, l2d, l3d, lat, lon, kx, kt, d, fl0 = random_case(seed, nr)
    go, gr = both(im, jm, nlev, nlev, plevs, table, l2d, l3d, nr)
    p1, p2 = ref.pint(nlev)
    lev = np.exp(rng.uniform(np.log(max(p2[-2] if nlev > 1 else 1.0, 0.5)), np.log(1999.0), len(lat)))
    lev[210:240] = p2[rng.integers(0, max(nlev - 1, 1), 30)]           # lev == an interface: belongs to the level above it
    lev[240:270] = p1[rng.integers(0, nlev, 30)]
    f1, f2 = fl0.copy(), fl0.copy()
    for rep in range(2):                                                  # two files into the same grids
        oracle.accum(go, nlev, lat, lon, lev, kx, kt, d, table, p1, p2, f1)
        ref.accum(gr, lat, lon, lev, kx, kt, d, table, p1, p2, f2)
        assert np.array_equal(f1, f2)
    same(go, gr, l2d, l3d, "raw sums")
    if nr == 3:
        oracle.norm_3ch(go, nlev, True, 0)
        ref.norm_3ch(gr, True, None)
        same(go, gr, l2d, l3d, "3CH Norm")
        return
    raw = {nm: getattr(go, nm).copy() for nm in go.names()}
    for nrmlz, ntresh in ((True, None), (False, None), (True, 0), (True, 1), (True, 3)):
        for g in (go, gr):
            for nm in go.names():
                getattr(g, nm)[...] = raw[nm]
        oracle.norm(go, nlev, nrmlz, ntresh or 0)
        ref.norm(gr, nrmlz, ntresh)
        if ntresh and ntresh > 1:
            # a 3-D box with exactly one observation below the threshold: the reference reads temporaries of an earlier
            # box there (m_gritas_grids.f:642-644) -- DESIGN.md section 5, known divergence; everything else must agree
            one = (raw["nobs_3d"] == 1.0)
            for nm in ("stdv_3d", "xcov_3d"):
                a, b = getattr(go, nm), getattr(gr, nm)
                m = np.broadcast_to(one[..., None], a.shape)
                a[m] = b[m]
        same(go, gr, l2d, l3d, f"Norm nrmlz={nrmlz} ntresh={ntresh}")


def test_level_miss_stops_at_the_same_statement():
    rng, im, jm, nlev, plevs, table, l2d, l3d, lat, lon, kx, kt, d, fl0 = random_case(11, 1, n=500, nlev=5)
    go, gr = both(im, jm, nlev, nlev, plevs, table, l2d, l3d, 1)
    p1, p2 = ref.pint(nlev)
    lev = np.full(len(lat), 500.0)
    lev[137] = 2500.0                                                     # above p1(1) = 2000 hPa: no level
    kx[137], kt[137], fl0[137] = 10, 44, 0                                # an upper-air variable, not rejected
    with pytest.raises(ref.RefError) as e:
        ref.accum(gr, lat, lon, lev, kx, kt, d, table, p1, p2, fl0.copy())
    assert e.value.line in (458, 459)                                     # 'Accum: could not find level'
    with pytest.raises(oracle.LevelMiss) as eo:
        oracle.accum(go, nlev, lat, lon, lev, kx, kt, d, table, p1, p2, fl0.copy())
    assert eo.value.args and float(eo.value.lev) == 2500.0
    same(go, gr, l2d, l3d, "sums up to the offending observation")       # both stop there: what came before is in the grids


def test_single_level():
    """nlevs = 1: p1 = p, p2 = p - 0.1 (:249-256); Norm returns before the 3-D block (:607)."""
    rng, im, jm, nlev, plevs, table, l2d, l3d, lat, lon, kx, kt, d, fl0 = random_case(5, 2, n=800, nlev=1)
    go, gr = both(im, jm, 1, 1, plevs, table, l2d, l3d, 2)
    p1, p2 = ref.pint(1)
    lev = np.full(len(lat), plevs[0] - 0.05)
    f1, f2 = fl0.copy(), fl0.copy()
    oracle.accum(go, 1, lat, lon, lev, kx, kt, d, table, p1, p2, f1)
    ref.accum(gr, lat, lon, lev, kx, kt, d, table, p1, p2, f2)
    assert np.array_equal(f1, f2)
    oracle.norm(go, 1, True, 0)
    ref.norm(gr, True, None)
    same(go, gr, l2d, l3d, "nlevs = 1")


@pytest.mark.parametrize("seed", range(3))
def test_minmax(seed):
    rng, im, jm, nlev, plevs, table, l2d, l3d, lat, lon, kx, kt, d, fl0 = random_case(100 + seed, 1)
    go, gr = both(im, jm, nlev, nlev, plevs, table, l2d, l3d, 1)
    p1, p2 = ref.pint(nlev)
    lev = np.exp(rng.uniform(np.log(max(p2[-2], 0.5)), np.log(1999.0), len(lat)))
    oracle.minmax_init(go)
    oracle.minmax_init(gr)
    f1, f2 = fl0.copy(), fl0.copy()
    oracle.minmax(go, nlev, lat, lon, lev, kx, kt, d, table, p1, p2, f1)
    ref.minmax(gr, lat, lon, lev, kx, kt, d, table, p1, p2, f2)
    assert np.array_equal(f1, f2)
    for nm in ("bias_2d", "stdv_2d", "nobs_2d", "bias_3d", "stdv_3d", "nobs_3d"):
        assert np.array_equal(getattr(go, nm), getattr(gr, nm)), nm


# ----------------------------------------------------------------------------------------------------------------------
# gritas_kxkt.f:10-119 (GrITAS_KxKt), gritas.f:562-704 (residual selection), :712-735 (QC / time / longitude filter), :738

def test_kxkt_table():
    for kt_list, kx_lists in ((KT_LIST, KX_LISTS),
                              ([44, 4, 1, 2, 44], [[120, 130], [220], [181, 187], [181], [130]]),
                              ([-5, 6, 3], [[1, 2, 3, 4], [40], [7]]),
                              ([], [])):
        kxmax = 256
        o = oracle.kxkt(64, kxmax, kt_list, kx_lists)
        r = ref.kxkt(64, kxmax, kt_list, kx_lists)
        for a, b in zip(o, r):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("kt_list,kx_lists,what", [([99], [[1]], "kt too large"), ([0], [[1]], "kt negative"),
                                                   ([44], [[300]], "kx too large"), ([44] * 23, [[1]] * 23, "l3d too large"),
                                                   ([1] * 21, [[1]] * 21, "l2d too large")])
def test_kxkt_errors(kt_list, kx_lists, what):
    with pytest.raises(ref.RefError):
        ref.kxkt(64, 256, kt_list, kx_lists)
    with pytest.raises(ValueError):
        oracle.kxkt(64, 256, kt_list, kx_lists)


@pytest.mark.parametrize("iopt", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                  28, 29, 30, 31, 32, 33])
def test_residual_selection(iopt):
    rng = np.random.default_rng(iopt)
    n = 777
    omf, oma, obs, xm = (rng.normal(0.2, 1.5, n) for _ in range(4))
    xvec = rng.uniform(-0.2, 2.0, n)                      # some non-positive: the where / elsewhere of options 32, 33
    qc = rng.integers(0, 3, n).astype(np.int32)
    for tch in (False, True):
        for self_ in (False, True):
            for scale_res in ((0, 1, 2, 3) if iopt <= 3 else (0,)):
                xv = np.abs(xvec) + 0.1 if scale_res == 1 else xvec
                want = oracle.select_del_all(iopt, omf, oma, obs, xv, xm, tch=tch, self_=self_, scale_res=scale_res)
                nr = want.shape[1]
                got, qc_after, nstat = ref.select(iopt, nr, omf, oma, obs, xv, xm, qc, tch=tch, self_=self_, scale_res=scale_res)
                if iopt in (17, 19):
                    got = ref.jo_scale(iopt, got)         # gritas.f:738 (the oracle's entry includes it)
                assert np.array_equal(got, want), (iopt, tch, self_, scale_res)
                assert np.array_equal(qc_after, qc)
                assert nstat == (1 if iopt in (16, 17, 18, 19) else None)


def test_residual_selection_xcov_copies_the_first_column():
    """gritas.f:573-574, 586-587: with -xcov / -pxcov the second column repeats the first"""
    rng = np.random.default_rng(3)
    omf, oma, obs, xvec, xm = (rng.normal(0, 1, 50) for _ in range(5))
    for iopt, src in ((1, omf), (3, oma)):
        got, _, _ = ref.select(iopt, 2, omf, oma, obs, xvec, xm, np.zeros(50, np.int32), xcov=True)
        assert np.array_equal(got[:, 0], src) and np.array_equal(got[:, 1], src)


@pytest.mark.parametrize("iopt,col", [(25, "omf"), (31, "oma")])
def test_meanres_selection(iopt, col):
    """gritas.f:653-664, 680-691: member minus ensemble mean where the observation values match, passive otherwise"""
    rng = np.random.default_rng(iopt)
    n = 300
    omf, oma, obs, xvec, xm = (rng.normal(0, 1, n) for _ in range(5))
    m_obs = obs.copy()
    m_obs[::7] += 1.0                                     # these do not match the mean's observations
    m_omf, m_oma = rng.normal(0, 1, n), rng.normal(0, 1, n)
    qc = np.zeros(n, np.int32)
    got, qc_after, _ = ref.select(iopt, 1, omf, oma, obs, xvec, xm, qc, meanres=True, m_obs=m_obs, m_omf=m_omf, m_oma=m_oma)
    match = obs == m_obs
    src, msrc = (omf, m_omf) if col == "omf" else (oma, m_oma)
    assert np.array_equal(got[match, 0], (msrc - src)[match])
    assert np.array_equal(qc_after, np.where(match, 0, 7))


def test_invalid_iopt_is_fatal():
    z = np.zeros(4)
    with pytest.raises(ref.RefError):
        ref.select(99, 1, z, z, z, z, z, np.zeros(4, np.int32))


@pytest.mark.parametrize("ioptn", [1, -1, 3])
@pytest.mark.parametrize("window", [(0, -1, -999.0, 999.0), (-60, 120, -999.0, 999.0), (0, -1, -30.0, 45.5), (-90, 90, 10.0, 20.0),
                                    (0, 0, -180.0, 180.0)])
def test_qc_time_longitude_filter(ioptn, window):
    rng = np.random.default_rng(abs(ioptn) * 31 + int(window[1]) + 7)
    n = 5000
    qcexcl = rng.choice(np.array([0, 0, 0, 7, 1, 2, 9], np.int32), n)
    time = rng.integers(-180, 181, n).astype(np.int32)
    lon = rng.uniform(-180, 180, n)
    lon[:8] = [-30.0, 45.5, 10.0, 20.0, -180.0, 180.0, 179.999, -179.999]
    time[:6] = [-60, 120, -90, 90, 0, 1]
    t1, t2, lona, lonb = window
    want = oracle.filter_flags(qcexcl, time, lon, ioptn=ioptn, t1=t1, t2=t2, lona=lona, lonb=lonb, x_passive=7, x_pre_bad=1)
    got = ref.filter_flags(qcexcl, time, lon, ioptn=ioptn, t1=t1, t2=t2, lona=lona, lonb=lonb, x_passive=7, x_pre_bad=1)
    assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------------------------------------------
# channel / level cross-covariances: GrITAS_XCov_Accum_ (:667-848), GrITAS_XCov_Norm_ (:859-1008), GrITAS_XCov_Prs_Accum_ (:1862-2059)

@pytest.mark.parametrize("tch", [False, True])
@pytest.mark.parametrize("two_grids", [False, True])
@pytest.mark.parametrize("seed", [1, 2])
def test_xcov_accum_and_norm(seed, two_grids, tch):
    from tests import xcov_data
    km, nr = 24, (3 if tch else 2)
    achan = np.array([2, 3, 5, 8, 9, 13, 20], np.int32)
    d = xcov_data.make_profiles(seed, 60, km, achan, nr, two_grids=two_grids)
    ref.init(4, 3, km, 1, np.array([500.0]), d["table"].shape[0], d["table"].shape[1])
    bo, xo, no = oracle.xcov_alloc(len(achan), d["l3d"], nr)
    br, xr, nr_ = oracle.xcov_alloc(len(achan), d["l3d"], nr)
    f1, f2 = d["flag"].copy(), d["flag"].copy()
    for rep in range(2):                                                  # two files into the same sums
        p1 = oracle.xcov_accum(km, achan, d["lev"], d["kx"], d["kt"], d["del_"], d["table"], f1, d["l3d"], bo, xo, no, tch)
        p2 = ref.xcov_accum(achan, d["lev"], d["kx"], d["kt"], d["del_"], d["table"], f2, d["l3d"], br, xr, nr_, tch)
        assert p1 == p2 and np.array_equal(f1, f2)
    for a, b in ((bo, br), (xo, xr), (no, nr_)):
        assert np.array_equal(a, b)
    raw = [a.copy() for a in (bo, xo, no)]
    for self_ in ((False, True) if tch else (False,)):
        for nrmlz, ntresh in ((True, None), (False, None), (True, 5)):
            for dst, src in zip((bo, xo, no, br, xr, nr_), raw + raw):
                dst[...] = src
            oracle.xcov_norm(nr, nrmlz, tch, self_, d["l3d"], bo, xo, no, ntresh or 0)
            ref.xcov_norm(nr, nrmlz, tch, self_, d["l3d"], br, xr, nr_, ntresh)
            for a, b, nm in ((bo, br, "bias"), (xo, xr, "xcov"), (no, nr_, "nobs")):
                assert np.array_equal(a, b, equal_nan=True), (nm, self_, nrmlz, ntresh)


def test_xcov_accum_incomplete_profiles_are_fatal():
    ref.init(4, 3, 24, 1, np.array([500.0]), 8, 8)
    b, x, n = oracle.xcov_alloc(3, 1, 2)
    lev = np.arange(1.0, 26.0)                                            # 25 observations, km = 24
    z = np.zeros(25, np.int32)
    with pytest.raises(ref.RefError):
        ref.xcov_accum(np.array([1, 2, 3], np.int32), lev, z + 1, z + 1, np.zeros((25, 2), order="F"), np.ones((8, 8), np.int32), z.copy(),
                       1, b, x, n, False)


@pytest.mark.parametrize("tch", [False, True])
@pytest.mark.parametrize("seed", [3, 4])
def test_xcov_prs_accum(seed, tch):
    rng = np.random.default_rng(seed)
    nlev, nr = 9, (3 if tch else 2)
    plevs = np.sort(rng.uniform(20.0, 1000.0, nlev))[::-1].copy()
    ref.init(4, 3, nlev, nlev, plevs, 4, 4)
    p1, p2 = ref.pint(nlev)
    # soundings: runs of equal ks, 3..12 levels each, some levels twice in a sounding.
    # A DEFECT OF THE REFERENCE shapes this input: its run-length count (:1961-1975) books the first observation of a run
    # to the run before it, so the LAST sounding of a call is allocated one element short and gathered past the end of
    # rlev / rdel (and a sounding index that comes back later overflows by a whole run): undefined behaviour, the sums of
    # that sounding depend on what the heap holds.  The oracle and the CUDA path compute what the loop intends.  To pin
    # everything else bit for bit the call ends with a one-observation sounding whose residuals are zero: whatever the
    # reference reads for it, it adds one count and nothing else -- in both implementations.
    ks, lev = [], []
    for s in range(40):
        m = int(rng.integers(3, 13))
        ks += [100 + s] * m
        lev += list(np.exp(rng.uniform(np.log(25.0), np.log(1500.0), m)))
    ks.append(999)
    lev.append(500.0)
    ks, lev = np.array(ks, np.int32), np.array(lev)
    d = np.asfortranarray(rng.normal(0.1, 1.0, (len(ks), nr)))
    d[-1, :] = 0.0
    bo, xo, no = oracle.xcov_alloc(nlev, 1, nr)
    br, xr, nr_ = oracle.xcov_alloc(nlev, 1, nr)
    for rep in range(2):
        q1 = oracle.xcov_prs_accum(nlev, lev, ks, d, p1, p2, 1, bo, xo, no, tch)
        q2 = ref.xcov_prs_accum(lev, ks, d, p1, p2, 1, br, xr, nr_, tch)
        assert q1 == q2
    for a, b, nm in ((bo, br, "bias"), (xo, xr, "xcov"), (no, nr_, "nobs")):
        assert np.array_equal(a, b), nm


def test_xcov_prs_accum_level_miss():
    plevs = np.array([850.0, 500.0, 250.0])
    ref.init(4, 3, 3, 3, plevs, 4, 4)
    p1, p2 = ref.pint(3)
    b, x, n = oracle.xcov_alloc(3, 1, 2)
    with pytest.raises(ref.RefError):
        ref.xcov_prs_accum(np.array([2500.0]), np.array([1], np.int32), np.zeros((1, 2), order="F"), p1, p2, 1, b, x, n, False)


# ----------------------------------------------------------------------------------------------------------------------
# scorecard: the latitude rows of the regions (init_, :136-138, 163-175), ave_nomiss (:1797-1826) under scores2_'s loop nest

@pytest.mark.parametrize("jm", [181, 91, 46, 721, 25])
def test_region_rows(jm):
    ref.init(8, jm, 1, 1, np.array([500.0]), 4, 4)
    assert np.array_equal(ref.score_regions(), oracle.score_regions(jm))


@pytest.mark.parametrize("seed", [0, 1])
def test_scorecard_reductions(seed):
    """scores2_ (:1671-1795) is a loop nest around ave_nomiss: surface variables first (level slot 1), then every level of
    every upper-air variable, for mean / rms / stdv and every region; collect_stats is real(4)."""
    rng = np.random.default_rng(seed)
    im, jm, km, l2d, l3d = 24, 19, 4, 2, 3
    ref.init(im, jm, km, km, np.array([850.0, 500.0, 250.0, 100.0]), 4, 4)
    g = oracle.Grids(im, jm, km, l2d, l3d, 1)
    for nm in g.names():
        a = getattr(g, nm)
        a[...] = rng.normal(0.0, 2.0, a.shape) if not nm.startswith("nobs") else rng.integers(0, 5, a.shape)
    for nm in ("bias", "rtms", "stdv"):                      # missing boxes
        for tag in ("2d", "3d"):
            a = getattr(g, f"{nm}_{tag}")
            a[rng.random(a.shape) < 0.3] = oracle.AMISS
    g.bias_3d[:, :, 2, 1, 0] = oracle.AMISS                  # a level without any valid box
    regindx = ref.score_regions()
    want = oracle.scores2(g, regindx)
    assert np.array_equal(ref.scores2(g, regindx), want)     # scores2_ itself, translated
    lon = np.array([1, im], np.int32)                        # ... and its loop nest spelled out around ave_nomiss
    got = np.zeros_like(want)
    for ne, nm in enumerate(("bias", "rtms", "stdv")):
        for mr in range(regindx.shape[1]):
            nv = 0
            for lm in range(l2d):
                got[mr, 0, nv, ne] = ref.ave_nomiss(getattr(g, nm + "_2d")[:, :, lm, 0], g.nobs_2d[:, :, lm], lon, regindx[:, mr])
                nv += 1
            for lm in range(l3d):
                for kk in range(km):
                    got[mr, kk, nv, ne] = ref.ave_nomiss(getattr(g, nm + "_3d")[:, :, kk, lm, 0], g.nobs_3d[:, :, kk, lm], lon,
                                                         regindx[:, mr])
                nv += 1
    assert np.array_equal(got, want)


# ----------------------------------------------------------------------------------------------------------------------
# land / ocean mask gather: maskout_ (m_gritas_masks.F90:137-186)

@pytest.mark.parametrize("seed", [0, 1, 2])
def test_maskout(seed):
    rng = np.random.default_rng(seed)
    im, jm = int(rng.integers(6, 80)), int(rng.integers(4, 50))
    n = 5000
    lat = rng.uniform(-95, 95, n)
    lon = rng.uniform(-200, 380, n)
    lon[:100] = -180.0 + (rng.integers(0, 2 * im, 100) + 0.5) * (360.0 / im)        # NINT ties, +-180
    lat[:100] = -90.0 + (rng.integers(0, 2 * jm, 100) * 0.5) * (180.0 / (jm - 1))
    mask = np.asfortranarray(rng.choice(np.array([0.0, 0.5, 0.98, 0.99, 0.995, 1.0]), (im, jm)))
    q0 = (rng.random(n) < 0.3).astype(np.int32)
    q1, q2 = q0.copy(), q0.copy()
    n1 = oracle.maskout(im, jm, lat, lon, mask, q1)
    n2 = ref.maskout(im, jm, lat, lon, mask, q2)
    assert n1 == n2 and np.array_equal(q1, q2)


# ----------------------------------------------------------------------------------------------------------------------
# the pre-sort (gritas.f:500-508): key sequence and direction from the source; the sort itself (m_MergeSorts, un-vendored) is a
# stable merge sort by assumption on both sides

@pytest.mark.parametrize("name", ["cfg1", "cfg5"])
def test_presort_key_sequence(name):
    cfg = synth.get_config(name, 0.01)
    c = synth.make_file(cfg, 1)
    a = oracle.presort(c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], c["ks"], c["kx"])
    b = ref.presort(c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], c["ks"], c["kx"])
    assert np.array_equal(a, b)
