#!/usr/bin/env python
"""Small invocations of every shared-memory / peer path of the tile kernel, for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck python tools/sanitize_smoke.py
tiled K1 + K2f, list overflow (direct RED + dirty tiles + K2t flush), export / import (multi-rank pipeline on one
rank), the sharded Norm with two handles (k_tile_final<TILE_PEERS>), the deterministic path and the pre-sort."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gritas_b200 import cabi, synth          # noqa: E402
from gritas_b200.m_gritas_grids import Grids  # noqa: E402


def run(mode, files, cfg, tables, env=None, sharded=0, pipeline=False):
    table, l2d, l3d, p1, p2 = tables
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        n = max(1, sharded)
        Gs = [Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode) for _ in range(n)]
        for G in Gs:
            G._ensure(2, table, p1, p2, l2d, l3d)
        if sharded:
            blobs = [G.peer_export() for G in Gs]
            for r, G in enumerate(Gs):
                G.peer_attach(n, r, blobs)
        for i, c in enumerate(files):
            Gs[i % n].accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        out = Gs[0].alloc_grids()
        if sharded:
            for G in Gs:
                G.sync_no_flush()
            for G in Gs:
                G.norm_sharded_into(out)
        elif pipeline:
            Gs[0].norm_reduce_into(out, root=0)
        else:
            Gs[0].norm_into(out)
            Gs[0].flush()
            Gs[0].norm_into(out)
        tot = float(out["nobs_3d"].sum())
        for G in Gs:
            G.close()
        return tot
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def main():
    cabi.load(build_if_missing=False)
    cfg = synth.get_config("cfg5", 0.01)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(3)]
    table, l2d, l3d = synth.kxkt_table(cfg)
    G0 = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs)
    p1, p2 = G0.GrITAS_Pint()
    tables = (table, l2d, l3d, p1, p2)
    T = cabi.MODE_FAST | cabi.FLAG_FORCE_TILED
    res = {
        "tiled": run(T, files, cfg, tables),
        "overflow": run(T, files, cfg, tables, env={"GRITAS_B200_STAGE_SLOTS": str(1 << 15)}),
        "export_import": run(T, files, cfg, tables, env={"GRITAS_B200_FORCE_PIPELINE": "1"}, pipeline=True),
        "sharded2": run(T, files, cfg, tables, sharded=2),
        "direct": run(cabi.MODE_FAST | cabi.FLAG_NO_STAGING, files, cfg, tables),
        "deterministic": run(cabi.MODE_DETERMINISTIC, files, cfg, tables),
    }
    c = files[0]
    ks = np.zeros(len(c["lat"]), np.int32)
    G0.presort(c["qcexcl"], c["lat"], c["lon"], c["lev"], c["time"], c["kt"], ks, c["kx"])
    vals = set(res.values())
    print(res)
    assert len(vals) == 1, "paths disagree on the number of binned observations"
    print("sanitize_smoke OK")


if __name__ == "__main__":
    main()
