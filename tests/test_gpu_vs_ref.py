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
