"""The CUDA path against the reference's OWN routines -- GrITAS_Pint_, GrITAS_Accum_, GrITAS_Norm_, GrITAS_MinMax_ of
m_gritas_grids.f, translated to C statement by statement (oracle/f2c_subset.py; the built oracle/_ref/libgritas_ref.so
travels to the GPU box).  Deterministic mode: flags, counts and every statistic BIT FOR BIT; fast mode: counts and flags
exact, statistics within 1e-6.  Skipped where the translated reference is not built."""
import numpy as np
import pytest

from gritas_b200 import cabi, synth
from gritas_b200.m_gritas_grids import Grids
from oracle import oracle, ref
from tests import helpers as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref is not built and /root/reference is not here")]


def ref_run(cfg, files, nr, fins, nrmlz=True, ntresh=None):
    table, l2d, l3d = synth.kxkt_table(cfg)
    ref.init(cfg.im, cfg.jm, cfg.km, cfg.nlevs, cfg.plevs, table.shape[0], table.shape[1])
    p1, p2 = ref.pint(cfg.nlevs)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    flags = []
    for c, fin in zip(files, fins):
        fl = fin.copy()
        ref.accum(g, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, nr), table, p1, p2, fl)
        flags.append(fl)
    ref.norm(g, nrmlz, ntresh)
    return g, flags, l2d, l3d


@pytest.mark.parametrize("mode", H.MODES)
@pytest.mark.parametrize("name,scale,nr", [("cfg1", 0.02, 1), ("cfg2", 0.02, 2), ("cfg5", 0.02, 1), ("cfg3_amsua", 0.002, 1)])
def test_cuda_equals_the_translated_reference(name, scale, nr, mode):
    cfg = synth.get_config(name, scale)
    if name != "cfg3_amsua":
        cfg.im, cfg.jm = 72, 46
    else:
        cfg.im, cfg.jm = 144, 73
    files = [synth.make_file(cfg, i) for i in range(3)]
    fins = [H.filter_in(c, 101) for c in files]
    g, rflags, l2d, l3d = ref_run(cfg, files, nr, fins)
    out, gflags = H.gpu_run(cfg, files, nr, mode, fins)
    for a, b in zip(gflags, rflags):
        assert np.array_equal(a, b)
    H.compare(out, g, l2d, l3d, nr, exact=H.is_det(mode), tol=1e-6)


def test_cuda_pint_equals_the_translated_reference():
    for nlev in (1, 2, 26, 72):
        rng = np.random.default_rng(nlev)
        plevs = np.sort(rng.uniform(0.5, 1050.0, nlev))[::-1].copy()
        ref.init(8, 5, nlev, nlev, plevs, 4, 4)
        r1, r2 = ref.pint(nlev)
        G = Grids(8, 5, nlev, plevs)
        p1, p2 = G.GrITAS_Pint()
        G.close()
        assert np.array_equal(p1, r1) and np.array_equal(p2, r2)


@pytest.mark.parametrize("nrmlz,ntresh", [(False, None), (True, 1)])
def test_cuda_norm_variants_equal_the_translated_reference(nrmlz, ntresh):
    cfg = synth.get_config("cfg2", 0.01)
    cfg.im, cfg.jm = 48, 25
    files = [synth.make_file(cfg, i) for i in range(2)]
    fins = [H.filter_in(c, 101) for c in files]
    g, rflags, l2d, l3d = ref_run(cfg, files, 2, fins, nrmlz, ntresh)
    out, gflags = H.gpu_run(cfg, files, 2, cabi.MODE_DETERMINISTIC, fins, nrmlz=nrmlz, ntresh=ntresh)
    H.compare(out, g, l2d, l3d, 2, exact=True, tol=0.0)


# ---- channel / level cross-covariances: the CUDA kernels (csrc/xcov.cu) against GrITAS_XCov_Accum_ / _Norm_ / _Prs_Accum_ as
# translated from m_gritas_grids.f:667-848, 859-1008, 1862-2059 ----

@pytest.mark.parametrize("tch,nr", [(False, 2), (True, 3)])
def test_cuda_xcov_equals_the_translated_reference(tch, nr):
    from gritas_b200.m_gritas_grids import XCov
    from tests import xcov_data as X
    km, achan = 24, [2, 3, 5, 7, 11, 12, 13, 20, 24]
    files = [X.make_profiles(10 + i, 70, km, achan, nr, two_grids=True) for i in range(2)]
    table, l3d = files[0]["table"], files[0]["l3d"]
    ref.init(4, 3, km, 1, np.array([500.0]), table.shape[0], table.shape[1])
    G = XCov(km, achan, l3d, nr, tch, table)
    b, x, n = oracle.xcov_alloc(len(achan), l3d, nr)
    for d in files:
        fr, fg = d["flag"].copy(), d["flag"].copy()
        np_r = ref.xcov_accum(np.array(achan, np.int32), d["lev"], d["kx"], d["kt"], d["del_"], table, fr, l3d, b, x, n, tch)
        np_g = G.GrITAS_XCov_Accum(None, None, d["lev"], d["kx"], d["kt"], d["del_"], len(d["lev"]), nr, table, fg, l3d)
        assert np_g == np_r and np.array_equal(fg, fr)
    gb, gx, gn = G.alloc()
    G.get_raw(gb, gx, gn)
    assert n.sum() > 0 and np.array_equal(gn, n) and np.array_equal(gb, b) and np.array_equal(gx, x)
    ref.xcov_norm(nr, True, tch, False, l3d, b, x, n, None)
    G.GrITAS_XCov_Norm(nr, True, tch, False, l3d, gb, gx, gn, None)
    assert np.array_equal(gb, b) and np.array_equal(gx, x)
    G.close()


@pytest.mark.parametrize("tch,nr", [(False, 2), (True, 3)])
def test_cuda_xcov_prs_equals_the_translated_reference(tch, nr):
    """The call ends with a one-observation sounding of zero residuals: the reference allocates the LAST sounding of a call one
    element short (m_gritas_grids.f:1961-1975) and reads past it; a zero sounding adds one count and nothing else either way."""
    from gritas_b200.m_gritas_grids import XCov
    plevs = np.array([1000, 925, 850, 700, 500, 400, 300, 250, 200, 150, 100, 70, 50, 30, 20, 10], dtype=np.float64)
    nl = len(plevs)
    ref.init(4, 3, nl, nl, plevs, 4, 4)
    p1, p2 = ref.pint(nl)
    G = XCov.for_levels(p1, p2, 1, nr, tch)
    b, x, n = oracle.xcov_alloc(nl, 1, nr)
    rng = np.random.default_rng(11)
    for f in range(2):
        ks_list, lev_list = [], []
        for s in range(200):
            m = int(rng.integers(1, 30))
            ks_list += [1000 + s] * m
            lev_list += list(np.exp(rng.uniform(np.log(8.0), np.log(1040.0), m)))
        ks_list.append(9999)
        lev_list.append(500.0)
        ks, lev = np.array(ks_list, np.int32), np.array(lev_list)
        d = np.asfortranarray(rng.normal(0.1, 1.0, (len(ks), nr)))
        d[-1, :] = 0.0
        np_r = ref.xcov_prs_accum(lev, ks, d, p1, p2, 1, b, x, n, tch)
        np_g = G.GrITAS_XCov_Prs_Accum(lev, None, None, ks, d, len(ks), nr)
        assert np_g == np_r
    gb, gx, gn = G.alloc()
    G.get_raw(gb, gx, gn)
    assert n.sum() > 0 and np.array_equal(gn, n) and np.array_equal(gb, b) and np.array_equal(gx, x)
    G.close()
