"""Host-side table builders of the library (GrITAS_Pint, GrITAS_KxKt) against the oracle."""
import numpy as np
import pytest

from gritas_b200 import synth
from gritas_b200.m_gritas_grids import Grids, GrITAS_KxKt, GritasPerror
from oracle import oracle, oracle_np


@pytest.mark.parametrize("plevs", [synth.PLEVS26, np.arange(616, 0, -1.0), np.arange(15, 0, -1.0), [500.0],
                                   [1000.0, 1.0], [1050.0, 1000.0, 999.99, 0.01]])
def test_pint_matches_oracle_bitwise(lib, plevs):
    G = Grids(8, 5, len(plevs), plevs)
    p1, p2 = G.GrITAS_Pint()
    o1, o2 = oracle.pint(plevs)
    assert np.array_equal(p1, o1) and np.array_equal(p2, o2)
    n1, n2 = oracle_np.pint(plevs)
    assert np.array_equal(p1, n1) and np.array_equal(p2, n2)
    if len(plevs) > 1:
        assert p1[0] == 2000.0 and p2[-1] == 0.0                # m_gritas_grids.f:270, 280
        assert np.array_equal(p1[1:], p2[:-1])                  # contiguous intervals
    else:
        assert p1[0] == plevs[0] and p2[0] == plevs[0] - 0.1    # :253-256


@pytest.mark.parametrize("plevs,msg", [([], "at least 1"), ([500.0, 700.0], "monotonically"),
                                       ([500.0, 500.0], "monotonically"), ([500.0, 0.0], "0 hPa")])
def test_pint_errors(lib, plevs, msg):
    G = Grids(8, 5, max(len(plevs), 1), plevs)
    with pytest.raises(GritasPerror, match=msg):
        G.GrITAS_Pint()
    with pytest.raises(ValueError, match=msg):
        oracle.pint(plevs)


def test_kxkt_matches_oracle(lib):
    for name in ("cfg1", "cfg3_iasi", "cfg4"):
        cfg = synth.get_config(name)
        lim3 = 22 if name != "cfg4" else 400          # cfg4 lifts gritas.h l3d_max=22
        t, l2, l3, n2, n3 = GrITAS_KxKt(cfg.ktmax, cfg.kxmax, 20, lim3, len(cfg.kt_list), cfg.kt_list, cfg.kx_lists)
        o = oracle.kxkt(cfg.ktmax, cfg.kxmax, cfg.kt_list, cfg.kx_lists, l3d_max=lim3)
        assert np.array_equal(t, o[0]) and (n2, n3) == (o[3], o[4])
        assert np.array_equal(l2, o[1]) and np.array_equal(l3, o[2])
        s = synth.kxkt_table(cfg)
        assert np.array_equal(t, s[0]) and (n2, n3) == (s[1], s[2])


def test_kxkt_semantics(lib):
    """surface kt (1,2,3 or negative) -> -l2d; upper air -> +l3d; last row wins (gritas_kxkt.f:105);
    multiple kx per row; terminator."""
    kt_list = [44, 4, 1, -7, 44]
    kx_lists = [[120, 130, 131], [220], [181, 187], [300], [130]]
    t, l2, l3, n2, n3 = GrITAS_KxKt(50, 400, 20, 22, 5, kt_list, kx_lists)
    assert (n2, n3) == (2, 3)
    assert list(l2) == [3, 4] and list(l3) == [1, 2, 5]
    assert t[119, 43] == 1 and t[130, 43] == 1
    assert t[129, 43] == 3                      # kx=130,kt=44 named by rows 1 and 5 -> row 5 (3rd 3-D grid)
    assert t[219, 3] == 2
    assert t[180, 0] == -1 and t[186, 0] == -1
    assert t[299, 6] == -2                      # kt=-7: surface by sign, stored under |kt|
    assert np.count_nonzero(t) == 7
    o = oracle.kxkt(50, 400, kt_list, kx_lists)
    assert np.array_equal(t, o[0])


@pytest.mark.parametrize("kt_list,kx_lists,lim,msg", [
    ([60], [[1]], (20, 22), "kt - too large"), ([0], [[1]], (20, 22), "kt - negative"),
    ([1, 2, 3], [[1], [2], [3]], (2, 22), "l2d - too large"), ([44, 45], [[1], [2]], (20, 1), "l3d - too large"),
    ([44], [[500]], (20, 22), "kx - too large")])
def test_kxkt_errors(lib, kt_list, kx_lists, lim, msg):
    with pytest.raises(GritasPerror, match=msg):
        GrITAS_KxKt(50, 400, lim[0], lim[1], len(kt_list), kt_list, kx_lists)
    with pytest.raises(ValueError, match=msg):
        oracle.kxkt(50, 400, kt_list, kx_lists, l2d_max=lim[0], l3d_max=lim[1])
