"""-meanres residual options (gritas.f:653-677, 680-691): member ODS against the ensemble-mean ODS.  Oracle: del and the
qcexcl change built on the host exactly as the reference does, then the filter and GrITAS_Accum_; the GPU entry does both on
the device.  Passive data are dropped (ioptn < 0), where the reference's del is defined."""
import numpy as np
import pytest

from gritas_b200 import cabi, synth
from gritas_b200.m_gritas_grids import Grids
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("iopt", [25, 27, 31])
@pytest.mark.parametrize("mode", [cabi.MODE_DETERMINISTIC, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED])
def test_meanres_matches_host_built_del(iopt, mode):
    cfg = synth.get_config("cfg1", 0.05)
    cfg.im, cfg.jm = 72, 46
    nr = 2 if iopt == 27 else 1
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, nr)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
    G._ensure(nr, table, p1, p2, l2d, l3d)
    for i in range(2):
        c = synth.make_file(cfg, i)
        rng = np.random.default_rng(40 + i)
        n = len(c["lat"])
        m = {"obs": c["obs"].copy(), "omf": c["omf"] + rng.normal(0, 0.2, n), "oma": c["oma"] + rng.normal(0, 0.1, n)}
        mism = rng.random(n) < 0.03
        m["obs"][mism] += 1.0                                   # not the same observation in the mean file
        qc = c["qcexcl"].copy()
        qc[mism] = synth.X_PASSIVE                              # gritas.f:660
        d = np.zeros((n, nr), order="F")
        d[:, 0] = (m["oma"] - c["oma"]) if iopt == 31 else (m["omf"] - c["omf"])
        if nr == 2:
            d[:, 1] = c["xvec"]
        fl = oracle.filter_flags(qc, c["time"], c["lon"], ioptn=-iopt, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        oracle.accum(ref, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
        fo = np.full(n, -3, np.int32)
        G.accum_odsx_meanres(c, m, iopt, ioptn=-iopt, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD, ob_flag_out=fo)
        assert np.array_equal(fo, fl)
    out = G.alloc_grids()
    G.get_raw(out["bias_2d"] if l2d else None, out["rtms_2d"] if l2d else None, out["xcov_2d"] if l2d else None,
              out["nobs_2d"] if l2d else None, out["bias_3d"], out["rtms_3d"], out["xcov_3d"], out["nobs_3d"])
    G.close()
    H.compare(out, ref, l2d, l3d, nr, exact=H.is_det(mode), tol=1e-6, raw=True)
