"""The compiled C++ host side (csrc/m_gritas_grids.hpp + gritas_driver.cpp, a stand-in for Program GrITAS,
gritas.f:389-808 / 885-887) drives the C ABI exactly as the Fortran caller would."""
import os
import subprocess

import numpy as np
import pytest

from gritas_b200 import build, synth
from oracle import oracle
from tests import helpers as H


def _prepare(tmp_path, nfiles=2):
    cfg = synth.get_config("cfg1", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    setup = os.path.join(tmp_path, "setup.bin")
    synth.write_setup(setup, cfg)
    paths = []
    for i, c in enumerate(files):
        p = os.path.join(tmp_path, f"t{i}.sods")
        synth.write_flat(p, c)
        paths.append(p)
    return cfg, files, setup, paths


def test_driver_builds_and_refuses_to_run_without_a_gpu(tmp_path, lib):
    exe = build.build_driver()
    assert os.path.exists(exe)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg, files, setup, paths = _prepare(str(tmp_path), 1)
    r = subprocess.run([exe, setup, os.path.join(tmp_path, "out.bin")] + paths, capture_output=True, text=True)
    assert r.returncode == 7                                   # gritas_perror -> exit(7): no CPU fallback
    assert "cannot create the B200 context" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("flags,iopt,ioptn,exact", [(["-omf"], 1, 1, False), (["-oma", "-det", "-nopassive"], 3, -3, True)])
def test_driver_matches_oracle(tmp_path, lib, flags, iopt, ioptn, exact):
    exe = build.build_driver()
    cfg, files, setup, paths = _prepare(str(tmp_path))
    out = os.path.join(tmp_path, "out.bin")
    r = subprocess.run([exe, setup, out] + flags + paths, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 1)
    for c in files:
        fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=ioptn, x_passive=7, x_pre_bad=1)
        d = oracle.select_del(iopt, c["omf"], c["oma"])
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], d, table, p1, p2, fl)
    oracle.norm(g, cfg.nlevs, True, 0)
    with open(out, "rb") as f:
        hd = np.fromfile(f, np.int32, 3)
        assert tuple(hd) == (l2d, l3d, 1)
        n3 = cfg.im * cfg.jm * cfg.km * l3d
        got = {nm: np.fromfile(f, np.float64, n3 if nm == "nobs" else n3).reshape(getattr(g, nm + "_3d").shape, order="F")
               for nm in ("bias", "rtms", "stdv", "xcov", "nobs")}
    assert np.array_equal(got["nobs"], g.nobs_3d)
    for nm in ("bias", "rtms", "stdv", "xcov"):
        ref = getattr(g, nm + "_3d")
        if exact:
            assert np.array_equal(got[nm], ref), nm
        else:
            ok = ref != oracle.AMISS
            assert np.array_equal(got[nm] == oracle.AMISS, ~ok)
            assert np.allclose(got[nm][ok], ref[ok], rtol=1e-9, atol=1e-12), nm
