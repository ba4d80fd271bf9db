"""A/B timing of library settings on the device-resident cfg2 month (one process, the month generated once).

    python tools/ab_bench.py [--files 124] [--steps 5] "NAME=VALUE,NAME=VALUE" "..." ...

Every positional argument is one variant: a comma-separated list of environment settings read by
gritas_b200_create (GRITAS_B200_TILE_SHIFT, GRITAS_B200_TF_BATCHES, ...); "" is the default build.
Prints accumulate / Norm / step times (CUDA events on the library stream) and checks the count grid.
A different libgritas_b200.so is a different process: GRITAS_B200_LIBDIR=... python tools/ab_bench.py ...
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="*", default=[""])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--files", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--concat", type=int, default=1, help="hand the library CONCAT files per Accum call (what batching the launches would give)")
    args = ap.parse_args()

    import torch

    import bench
    from gritas_b200 import cabi, synth
    from gritas_b200.m_gritas_grids import Grids

    cfg = synth.get_config(args.config, 1.0)
    nfiles = args.files or cfg.nfiles
    files = bench.gen_files_idx(cfg, list(range(nfiles)))
    if args.concat > 1:
        files = [{k: np.concatenate([c[k] for c in files[i:i + args.concat]]) for k in files[0]} for i in range(0, nfiles, args.concat)]
    table, l2d, l3d = synth.kxkt_table(cfg)
    expected = int(sum(np.count_nonzero(np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)) &
                                        (table[c["kx"] - 1, c["kt"] - 1] != 0)) for c in files))
    nobs = sum(len(c["lat"]) for c in files)
    torch.cuda.set_device(0)
    dev = [{k: torch.from_numpy(v).cuda() for k, v in c.items()} for c in files]
    L = cabi.load()
    NR = cfg.nr
    IOPT = cabi.IOPT_OMF_OMA if NR == 2 else cabi.IOPT_OMF
    print(f"# {cfg.name}: {nfiles} files, {nobs} obs, {expected} accepted; lib {os.environ.get('GRITAS_B200_LIBDIR', 'default')}", flush=True)
    for var in args.variants:
        sets = [kv.split("=", 1) for kv in var.split(",") if kv]
        old = {k: os.environ.get(k) for k, _ in sets}
        for k, v in sets:
            os.environ[k] = v
        try:
            mode = (cabi.MODE_FAST if args.mode == "fast" else cabi.MODE_DETERMINISTIC) | cabi.FLAG_ASYNC | cabi.FLAG_HOLD_INPUTS
            G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode, device=0)
            p1, p2 = G.GrITAS_Pint()
            G._ensure(NR, table, p1, p2, l2d, l3d)
            stream = torch.cuda.ExternalStream(G.stream(), device=torch.device("cuda", 0))
            calls = []
            for c in dev:
                p = lambda k: cabi.ptr(c[k])
                calls.append((G.handle, int(c["lat"].shape[0]), p("lat"), p("lon"), p("lev"), None, p("kt"), p("kx"),
                              p("qcexcl"), p("omf"), p("oma") if NR == 2 else None, None, None, IOPT, 0, IOPT, 0, -1,
                              -999.0, 999.0, synth.X_PASSIVE, synth.X_PRE_BAD, None, None))
            dptr = [bench.C_void() for _ in range(10)]

            def step(marks=None):
                G.reset()
                if marks:
                    marks[0].record(stream)
                for a in calls:
                    rc = L.gritas_b200_accum_ods(*a)
                    if rc:
                        raise RuntimeError(f"accum_ods rc={rc}: {G._err()}")
                G.barrier()
                if marks:
                    marks[1].record(stream)
                rc = L.gritas_b200_norm_device(G.handle, 1, -1, *[None] * 5, *dptr[5:])
                if rc:
                    raise RuntimeError(f"norm_device rc={rc}: {G._err()}")
                if marks:
                    marks[2].record(stream)

            for _ in range(args.warmup):
                step()
            torch.cuda.synchronize()
            ev = lambda: torch.cuda.Event(enable_timing=True)
            marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
            for m in marks:
                step(m)
            torch.cuda.synchronize()
            acc = sorted(m[0].elapsed_time(m[1]) for m in marks)
            nrm = sorted(m[1].elapsed_time(m[2]) for m in marks)
            tot = sorted(m[0].elapsed_time(m[2]) for m in marks)
            out = G.alloc_grids()
            G.norm_into(out)
            got = int(out["nobs_3d"].sum() + out["nobs_2d"].sum())
            med = lambda x: x[len(x) // 2]
            print(f"{var or 'default':48s} accum {med(acc):7.3f}  norm {med(nrm):7.3f}  step {med(tot):7.3f} ms  (min {tot[0]:.3f})  "
                  f"{nobs / med(tot) / 1e6:7.2f} G obs/s  counts {'OK' if got == expected else f'MISMATCH {got} != {expected}'}", flush=True)
            G.close()
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


if __name__ == "__main__":
    main()
