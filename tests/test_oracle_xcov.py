"""The CPU oracle's restatement of the cross-covariance family (m_gritas_grids.f:667-1008, 1862-2059) against an
independently written numpy formulation and hand-computed known answers."""
import numpy as np
import pytest

from oracle import oracle
from tests import xcov_data as X


@pytest.mark.parametrize("tch,nr", [(False, 2), (True, 3)])
@pytest.mark.parametrize("two", [False, True])
def test_xcov_accum_matches_numpy_twin(tch, nr, two):
    km, achan = 24, [2, 3, 5, 7, 11, 12, 13, 20, 24]
    d = X.make_profiles(7, 60, km, achan, nr, two_grids=two)
    b, x, n = oracle.xcov_alloc(len(achan), d["l3d"], nr)
    fl = d["flag"].copy()
    nprof = oracle.xcov_accum(km, achan, d["lev"], d["kx"], d["kt"], d["del_"], d["table"], fl, d["l3d"], b, x, n, tch)
    rb, rx, rn, rfl, racc = X.reference_numpy(km, achan, d, nr, tch)
    assert nprof == racc and 0 < nprof < 60
    assert np.array_equal(fl, rfl)
    assert np.array_equal(n, rn)
    assert np.allclose(b, rb, rtol=1e-13, atol=1e-13)
    assert np.allclose(x, rx, rtol=1e-12, atol=1e-12)


def test_xcov_accum_errors():
    km, achan = 8, [1, 2, 3]
    d = X.make_profiles(1, 4, km, achan, 2)
    b, x, n = oracle.xcov_alloc(3, 1, 2)
    with pytest.raises(ValueError, match="not all profiles complete"):
        oracle.xcov_accum(km, achan, d["lev"][:-1], d["kx"][:-1], d["kt"][:-1], d["del_"][:-1], d["table"], d["flag"][:-1].copy(), 1, b, x, n, False)
    with pytest.raises(ValueError, match="TCH setting"):
        oracle.xcov_accum(km, achan, d["lev"], d["kx"], d["kt"], d["del_"], d["table"], d["flag"].copy(), 1, b, x, n, True)
    lev = d["lev"].copy()
    lev[4] = 1.0                                    # channel 1 twice in the first profile: a fourth match
    lev[5] = 2.0
    with pytest.raises(ValueError, match="channel matching"):
        oracle.xcov_accum(km, achan, lev, d["kx"], d["kt"], d["del_"], d["table"], np.zeros(len(lev), np.int32), 1, b, x, n, False)


def test_xcov_norm_known_answers():
    """Two channels, three profiles, by hand (m_gritas_grids.f:938-989): Desroziers, symmetrised mean product."""
    x1 = np.array([[1.0, 2.0], [3.0, 5.0], [2.0, 2.0]])        # series 1, profiles x channels
    x2 = np.array([[0.5, 1.0], [1.0, 3.0], [4.0, 1.0]])        # series 2
    b, x, n = oracle.xcov_alloc(2, 1, 2)
    for p in range(3):
        b[:, 0, 0] += x1[p]
        b[:, 0, 1] += x2[p]
        x[:, :, 0, 0] += np.outer(x1[p], x2[p])
        n[:, 0] += 1.0
    oracle.xcov_norm(2, True, False, False, 1, b, x, n)
    m = 3
    mean1, mean2 = x1.sum(0) / m, x2.sum(0) / m
    assert np.array_equal(b[:, 0, 0], mean1) and np.array_equal(b[:, 0, 1], mean2)
    for k1 in range(2):
        for k2 in range(2):
            raw = sum(x1[p, k1] * x2[p, k2] for p in range(3))
            xt = raw / m - 0.5 * (mean1[k1] * mean2[k2] + mean1[k2] * mean2[k1])
            assert x[k1, k2, 0, 0] == m * xt / (m - 1)
    # a channel seen once is left un-normalised (m > 1 is required, :946, :974)
    b, x, n = oracle.xcov_alloc(2, 1, 2)
    b[:, 0, :] = 7.0
    x[:, :, 0, 0] = 3.0
    n[:, 0] = [1.0, 2.0]
    oracle.xcov_norm(2, True, False, False, 1, b, x, n)
    assert b[0, 0, 0] == 7.0 and b[1, 0, 0] == 3.5
    assert x[0, 0, 0, 0] == 3.0 and x[1, 0, 0, 0] == 3.0        # column k2 = 1 has n = 1: untouched
    assert x[0, 1, 0, 0] == 2 * (3.0 / 2 - 0.5 * (7.0 * 3.5 + 3.5 * 7.0)) / 1


def test_xcov_norm_three_cornered_hat():
    rng = np.random.default_rng(3)
    b, x, n = oracle.xcov_alloc(4, 1, 3)
    r = rng.normal(size=(50, 4, 3))
    for p in range(50):
        b[:, 0, :] += r[p]
        for s in range(3):
            x[:, :, 0, s] += np.outer(r[p, :, s], r[p, :, s])
        n[:, 0] += 1.0
    bs, xs = b.copy(order="F"), x.copy(order="F")
    oracle.xcov_norm(3, True, True, True, 1, bs, xs, n)           # self: the three variances
    oracle.xcov_norm(3, True, True, False, 1, b, x, n)
    cov = [np.cov(r[:, :, s].T, ddof=1) for s in range(3)]
    for s in range(3):
        assert np.allclose(xs[:, :, 0, s], cov[s], rtol=1e-11, atol=1e-12)
    assert np.array_equal(x[:, :, 0, 0], 0.5 * (xs[:, :, 0, 0] + xs[:, :, 0, 1] - xs[:, :, 0, 2]))
    assert np.array_equal(x[:, :, 0, 1], 0.5 * (xs[:, :, 0, 2] + xs[:, :, 0, 0] - xs[:, :, 0, 1]))
    assert np.array_equal(x[:, :, 0, 2], 0.5 * (xs[:, :, 0, 1] + xs[:, :, 0, 2] - xs[:, :, 0, 0]))


def test_xcov_prs_accum_by_hand():
    """Two soundings, levels found with GrITAS_Pint intervals (m_gritas_grids.f:2004-2013), l = 1."""
    plevs = np.array([1000.0, 500.0, 100.0])
    p1, p2 = oracle.pint(plevs)
    lev = np.array([990.0, 480.0, 110.0, 505.0, 950.0])
    ks = np.array([7, 7, 7, 9, 9], np.int32)
    d = np.asfortranarray(np.array([[1.0, 2.0], [3.0, 1.0], [2.0, 2.0], [4.0, 0.5], [1.0, 1.0]]))
    b, x, n = oracle.xcov_alloc(3, 1, 2)
    nprof = oracle.xcov_prs_accum(3, lev, ks, d, p1, p2, 1, b, x, n, False)
    assert nprof == 2
    kidx = [0, 1, 2, 1, 0]
    ex = np.zeros((3, 3))
    for grp in ([0, 1, 2], [3, 4]):
        for a in grp:
            for c in grp:
                ex[kidx[a], kidx[c]] += 0.5 * (d[a, 0] * d[c, 1] + d[a, 1] * d[c, 0])
    assert np.array_equal(x[:, :, 0, 0], ex)
    assert np.array_equal(n[:, 0], [2.0, 2.0, 1.0])
    assert np.array_equal(b[:, 0, 0], [2.0, 7.0, 2.0])
    with pytest.raises(oracle.LevelMiss):
        oracle.xcov_prs_accum(3, np.array([2500.0]), np.array([1], np.int32), d[:1], p1, p2, 1, b, x, n, False)
