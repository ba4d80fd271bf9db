"""Scorecard reductions (scores2_ + ave_nomiss + the region rows of init_, m_gritas_grids.f:163-175, 1671-1826):
oracle known answers on the CPU, GPU against the oracle."""
import numpy as np
import pytest

from gritas_b200 import cabi, synth
from gritas_b200.m_gritas_grids import Grids
from oracle import oracle
from tests import helpers as H


def test_score_regions_known_rows():
    # 1-degree grid, jm = 181: glat(j) = -90 + (j-1); global = rows 1..181, n.hem 20..80 -> rows 111..171, etc.
    r = oracle.score_regions(181)
    assert r[:, 0].tolist() == [1, 181]
    assert r[:, 1].tolist() == [111, 171]
    assert r[:, 2].tolist() == [71, 111]
    assert r[:, 3].tolist() == [11, 71]
    # the library's host helper is the same function
    G = Grids(360, 181, 26, synth.PLEVS26)
    assert np.array_equal(G.score_regions(), r)
    assert np.array_equal(Grids(72, 46, 26, synth.PLEVS26).score_regions(), oracle.score_regions(46))


def test_ave_nomiss_known_answers():
    im, jm, km = 4, 5, 2
    g = oracle.Grids(im, jm, km, 1, 1, 1)
    for a in (g.bias_2d, g.rtms_2d, g.stdv_2d, g.bias_3d, g.rtms_3d, g.stdv_3d):
        a[...] = oracle.AMISS
    g.bias_3d[0, 1, 0, 0, 0] = 2.0
    g.bias_3d[3, 2, 0, 0, 0] = 4.0
    g.bias_3d[1, 4, 0, 0, 0] = 100.0            # outside the region rows 2..3
    g.rtms_2d[2, 2, 0, 0] = 7.0
    reg = np.asfortranarray(np.array([[2, 3], [1, 5]], np.int32).T)
    out = oracle.scores2(g, reg)
    assert out.shape == (2, km, 2, 3)
    assert out[0, 0, 1, 0] == np.float32(3.0)                      # mean of 2 and 4
    assert out[1, 0, 1, 0] == np.float32(106.0 / 3.0)
    assert out[0, 1, 1, 0] == np.float32(oracle.AMISS)             # nothing valid on level 2
    assert out[0, 0, 0, 1] == np.float32(7.0) and out[0, 1, 0, 1] == 0.0   # surface: level slot 1 only
    assert out[0, 0, 0, 0] == np.float32(oracle.AMISS)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", H.MODES)
def test_gpu_scores_match_oracle(mode):
    cfg = synth.get_config("cfg1", 0.05)
    cfg.im, cfg.jm = 72, 46
    cfg.kt_list = [44, 4, 5, 1, 2]
    cfg.kx_lists = [[120, 130], [220], [220], [181, 187], [181]]
    rng = np.random.default_rng(5)
    files = []
    for i in range(2):
        c = synth.make_file(synth.get_config("cfg1", 0.05), i)
        row = rng.integers(0, 5, len(c["lat"]))
        c["kt"] = np.asarray(cfg.kt_list, np.int32)[row]
        c["kx"] = np.asarray([120, 220, 220, 181, 181], np.int32)[row]
        files.append(c)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 1)
    want = oracle.scores2(ref, oracle.score_regions(cfg.jm))
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    for c in files:
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), len(c["lat"]), 1, table, p1, p2, None, l2d, l3d=l3d)
    got = G.GrITAS_Scores()
    again = G.GrITAS_Scores()                       # the accumulation state is unchanged
    G.close()
    assert got.shape == want.shape == (4, cfg.km, l2d + l3d, 3)
    assert np.array_equal(got == np.float32(oracle.AMISS), want == np.float32(oracle.AMISS))
    if H.is_det(mode):
        assert np.array_equal(got, want)
    else:
        assert np.allclose(got, want, rtol=2e-6, atol=1e-7)
    assert np.allclose(got, again, rtol=2e-6, atol=1e-7)
