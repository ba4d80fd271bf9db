"""GrITAS_XCov_Accum_ / GrITAS_XCov_Norm_ (m_gritas_grids.f:667-1008) on the GPU against the CPU oracle: the sums are
formed profile by profile in input order with one multiply and one add per element, so everything is BIT-EXACT."""
import numpy as np
import pytest

from gritas_b200.m_gritas_grids import GritasPerror, XCov
from oracle import oracle
from tests import xcov_data as X

pytestmark = pytest.mark.gpu


def _both(km, achan, files, nr, tch, l3d, table, device_inputs=False):
    G = XCov(km, achan, l3d, nr, tch, table)
    b, x, n = oracle.xcov_alloc(len(achan), l3d, nr)
    keep = []
    for d in files:
        fo, fg = d["flag"].copy(), d["flag"].copy()
        np_o = oracle.xcov_accum(km, achan, d["lev"], d["kx"], d["kt"], d["del_"], table, fo, l3d, b, x, n, tch)
        args = [d["lev"], d["kx"], d["kt"], d["del_"], fg]
        if device_inputs:
            import torch
            args = [torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda() for a in args]
            keep.append(args)
        np_g = G.GrITAS_XCov_Accum(None, None, args[0], args[1], args[2], args[3], len(d["lev"]), nr, table, args[4], l3d)
        if device_inputs:
            fg = args[4].cpu().numpy()
        assert np_g == np_o
        assert np.array_equal(fg, fo)
    return G, (b, x, n)


@pytest.mark.parametrize("tch,nr", [(False, 2), (True, 3)])
@pytest.mark.parametrize("two", [False, True])
def test_xcov_accum_and_norm_bit_exact(tch, nr, two):
    km, achan = 24, [2, 3, 5, 7, 11, 12, 13, 20, 24]
    files = [X.make_profiles(10 + i, 70, km, achan, nr, two_grids=two) for i in range(3)]
    G, (b, x, n) = _both(km, achan, files, nr, tch, files[0]["l3d"], files[0]["table"])
    gb, gx, gn = G.alloc()
    G.get_raw(gb, gx, gn)
    assert np.array_equal(gn, n) and np.array_equal(gb, b) and np.array_equal(gx, x)
    for nrmlz, self_, ntresh in ((True, False, None), (True, True, 5), (False, False, None)):
        rb, rx = b.copy(order="F"), x.copy(order="F")
        oracle.xcov_norm(nr, nrmlz, tch, self_, files[0]["l3d"], rb, rx, n, 0 if ntresh is None else ntresh)
        G.GrITAS_XCov_Norm(nr, nrmlz, tch, self_, files[0]["l3d"], gb, gx, gn, ntresh)
        assert np.array_equal(gn, n)
        assert np.array_equal(gb, rb), (nrmlz, self_, ntresh)
        assert np.array_equal(gx, rx), (nrmlz, self_, ntresh)
    assert G.launches() > 0
    G.close()


def test_xcov_hyperspectral_size_device_inputs():
    """616 channels, 180 active (the IASI subset size of etc/gritas_iasi_metop-b.rc), device-resident columns; tiles of 64
    do not divide 180: ragged edges."""
    km = 616
    rng = np.random.default_rng(0)
    achan = np.sort(rng.choice(np.arange(1, km + 1), 180, replace=False)).tolist()
    files = [X.make_profiles(30 + i, 40, km, achan, 2) for i in range(2)]
    G, (b, x, n) = _both(km, achan, files, 2, False, 1, files[0]["table"], device_inputs=True)
    gb, gx, gn = G.alloc()
    G.get_raw(gb, gx, gn)
    assert n.max() > 0
    assert np.array_equal(gn, n) and np.array_equal(gb, b) and np.array_equal(gx, x)
    G.close()


def test_xcov_errors_like_the_reference():
    km, achan = 8, [1, 2, 3]
    d = X.make_profiles(1, 4, km, achan, 2)
    with pytest.raises(GritasPerror):
        XCov(km, achan, 1, 3, False, d["table"])              # Desroziers with three series (:742-744)
    G = XCov(km, achan, 1, 2, False, d["table"])
    with pytest.raises(GritasPerror, match="not all profiles complete"):
        G.GrITAS_XCov_Accum(None, None, d["lev"][:-1], d["kx"][:-1], d["kt"][:-1], np.asfortranarray(d["del_"][:-1]), len(d["lev"]) - 1, 2,
                            d["table"], d["flag"][:-1].copy(), 1)
    lev = d["lev"].copy()
    lev[4], lev[5] = 1.0, 2.0
    with pytest.raises(GritasPerror, match="channel matching"):
        G.GrITAS_XCov_Accum(None, None, lev, d["kx"], d["kt"], d["del_"], len(lev), 2, d["table"], np.zeros(len(lev), np.int32), 1)
    G.close()


@pytest.mark.parametrize("tch,nr", [(False, 2), (True, 3)])
def test_xcov_prs_accum_bit_exact(tch, nr):
    """GrITAS_XCov_Prs_Accum_ (m_gritas_grids.f:1862-2059): soundings by ks, several observations of a sounding on the same level,
    two files accumulated one after the other, and a sounding index that comes back later in the file (processed once per run)."""
    plevs = np.array([1000, 925, 850, 700, 500, 400, 300, 250, 200, 150, 100, 70, 50, 30, 20, 10], dtype=np.float64)
    p1, p2 = oracle.pint(plevs)
    nl = len(plevs)
    G = XCov.for_levels(p1, p2, 1, nr, tch)
    b, x, n = oracle.xcov_alloc(nl, 1, nr)
    rng = np.random.default_rng(11)
    for f in range(2):
        ks_list, lev_list = [], []
        for s in range(300):
            m = int(rng.integers(1, 40))
            ks_list += [1000 + s] * m
            lev_list += list(np.exp(rng.uniform(np.log(8.0), np.log(1040.0), m)))
        if f == 1:                                   # sounding 1005 again at the end of the file
            ks_list += [1005] * 7
            lev_list += list(np.exp(rng.uniform(np.log(8.0), np.log(1040.0), 7)))
        ks = np.array(ks_list, np.int32)
        lev = np.array(lev_list)
        d = np.asfortranarray(rng.normal(0.1, 1.0, (len(ks), nr)))
        np_o = oracle.xcov_prs_accum(nl, lev, ks, d, p1, p2, 1, b, x, n, tch)
        np_g = G.GrITAS_XCov_Prs_Accum(lev, None, None, ks, d, len(ks), nr)
        assert np_g == np_o
    gb, gx, gn = G.alloc()
    G.get_raw(gb, gx, gn)
    assert n.sum() > 0
    assert np.array_equal(gn, n) and np.array_equal(gb, b) and np.array_equal(gx, x)
    rb, rx = b.copy(order="F"), x.copy(order="F")
    oracle.xcov_norm(nr, True, tch, False, 1, rb, rx, n)
    G.GrITAS_XCov_Norm(nr, True, tch, False, 1, gb, gx, gn)
    assert np.array_equal(gb, rb) and np.array_equal(gx, rx)
    with pytest.raises(GritasPerror, match="could not find level"):
        G.GrITAS_XCov_Prs_Accum(np.array([500.0, 2500.0]), None, None, np.array([1, 1], np.int32), np.zeros((2, nr), order="F"), 2, nr)
    G.close()
