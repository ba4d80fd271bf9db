#!/usr/bin/env python
"""Where does the drop-in's pageable end-to-end time go?  Times, for one month of cfg2 columns in pageable numpy arrays:
the GrITAS_Accum loop (synchronous with ob_flag write-back / ASYNC without write-back) and GrITAS_Norm into pageable
and into pinned arrays.  GRITAS_B200_COPY_THREADS sets the size of the library's host copy pool (per process)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gritas_b200 import cabi, synth          # noqa: E402
from gritas_b200.m_gritas_grids import Grids  # noqa: E402


def main():
    nfiles = int(sys.argv[1]) if len(sys.argv) > 1 else 31
    cfg = synth.get_config("cfg2", 1.0)
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    table, l2d, l3d = synth.kxkt_table(cfg)
    L = cabi.load()
    res = {"copy_threads": os.environ.get("GRITAS_B200_COPY_THREADS", "default"), "files": nfiles}
    for name, mode in (("sync_writeback", cabi.MODE_FAST), ("async_nowriteback", cabi.MODE_FAST | cabi.FLAG_ASYNC | cabi.FLAG_NO_WRITEBACK)):
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
        p1, p2 = G.GrITAS_Pint()
        G._ensure(2, table, p1, p2, l2d, l3d)
        calls = []
        keep = []
        for c in files:
            d = np.empty((len(c["lat"]), 2), order="F")
            d[:, 0], d[:, 1] = c["omf"], c["oma"]
            fl = np.where(np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)), 0, 1).astype(np.int32)
            keep.append((d, fl))
            bad = C.c_double(0.0)
            calls.append((G.handle, len(fl), cabi.ptr(c["lat"]), cabi.ptr(c["lon"]), cabi.ptr(c["lev"]), cabi.ptr(d), cabi.ptr(c["kx"]),
                          cabi.ptr(c["kt"]), cabi.ptr(fl), C.byref(bad)))
        pout = G.alloc_grids()
        hout = G.alloc_grids(pinned=True)

        def accum():
            G.reset()
            for a in calls:
                rc = L.gritas_b200_accum(*a)
                assert rc == 0, rc
            G.sync_no_flush()
        for rep in range(2):
            t0 = time.perf_counter(); accum(); t1 = time.perf_counter()
            G.norm_into(pout); t2 = time.perf_counter()
            G.norm_into(hout); t3 = time.perf_counter()
        nobs = sum(len(c["lat"]) for c in files)
        res[name] = {"accum_ms_per_file": 1e3 * (t1 - t0) / nfiles, "accum_gbs": nobs * 56 / (t1 - t0) / 1e9,
                     "norm_pageable_ms": 1e3 * (t2 - t1), "norm_pinned_ms": 1e3 * (t3 - t2)}
        G.close()
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
