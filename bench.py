#!/usr/bin/env python
"""bench.py -- observations binned per second on the BASELINE.json workload.

A "step" is one pass of the hot path over one month of synthetic conventional ODS columns
(BASELINE.json configs[1]: 124 six-hourly files x ~1.21 M obs, omf+oma in one pass, 1x1 deg x 26
levels, 16 (kt,kx) grids): zero the accumulators, accumulate every file (QC filter + box index +
count/sum/sum-of-squares, fused kernel), [all-reduce the partial grids over ranks], finalize
mean / rms / stdv.  With N ranks every rank processes its own month (weak scaling) and the grids
are summed with one NCCL all-reduce before the finalize.

  value  : whole-job obs/s with the columns already resident in HBM (CUDA events on the library stream)
  e2e    : same step through the C ABI with pinned HOST columns (H2D inside) and the finalized
           grids copied back to pinned host arrays (D2H inside)
  --impl reference : the CPU restatement of the reference loops (oracle/, the reference itself cannot
           be built here: Fortran + un-vendored GMAO_Shared), file-sharded over the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "observations_binned_per_sec"
UNIT = "obs/s"
COLS_F64 = ("lat", "lon", "lev", "omf", "oma")
COLS_I32 = ("kt", "kx", "qcexcl")
def b_obs(nr):
    """algorithmic bytes per observation (SURVEY.md 8d / BASELINE.md 3): 44 B at nr=1, 52 B at nr=2"""
    return 36 + 8 * nr


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--files", type=int, default=0, help="override the number of files (debug)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink obs per file (debug)")
    ap.add_argument("--mode", default="fast", choices=["fast", "deterministic"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-staging", action="store_true", help="fast mode without the tiled accumulate")
    return ap.parse_args()


def gen_files(cfg, first, count, workers=8):
    from gritas_b200 import synth

    def one(i):
        c = synth.make_file(cfg, first + i)
        return {k: np.ascontiguousarray(c[k]) for k in COLS_F64 + COLS_I32}
    with ThreadPoolExecutor(max_workers=max(1, min(workers, os.cpu_count() or 1))) as ex:
        return list(ex.map(one, range(count)))


def workload_name(cfg, nfiles):
    what = "omf+oma (nr=2)" if cfg.nr == 2 else "omf (nr=1)"
    return (f"{cfg.name}: {nfiles} files x {cfg.nobs_per_file} obs, {what}, {cfg.im}x{cfg.jm}x{cfg.km}, "
            f"{len(cfg.kt_list)} (kt,kx) grids")


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_ev = threading.Event()
        self.ready = threading.Event()
        self.active = False            # only samples taken while the timed region runs are kept
        self.sm, self.reasons, self.max = [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, n): n[len("nvmlClocksEventReason"):] for n in dir(nv)
                     if n.startswith("nvmlClocksEventReason") and isinstance(getattr(nv, n), int)}
            if not names:
                names = {getattr(nv, n): n[len("nvmlClocksThrottleReason"):] for n in dir(nv)
                         if n.startswith("nvmlClocksThrottleReason") and isinstance(getattr(nv, n), int)}
            self.ready.set()
            while not self.stop_ev.is_set():
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.active:
                    self.sm.append(clk)
                    for bit, nm in names.items():
                        if bit and (r & bit) and nm not in ("None", "All", "GpuIdle", "ApplicationsClocksSetting"):
                            self.reasons.add(nm)
                time.sleep(0.005)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"sampler_error:{type(e).__name__}")
            self.ready.set()

    def result(self):
        self.stop_ev.set()
        self.join(timeout=2)
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(cfg, files, table, l2d, l3d, p1, p2, nthreads, grids):
    """One CPU step: the files sharded over `nthreads` workers (private grids each, like independent
    gritas.x processes), QC filter + GrITAS_Accum per file, partial grids summed, GrITAS_Norm."""
    from gritas_b200 import synth
    from oracle import oracle

    def work(t):
        g = grids[t]
        for a in (g.bias_3d, g.rtms_3d, g.xcov_3d, g.nobs_3d, g.bias_2d, g.rtms_2d, g.xcov_2d, g.nobs_2d):
            a[...] = 0.0
        for c in files[t::nthreads]:
            fl = oracle.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE,
                                     x_pre_bad=synth.X_PRE_BAD)
            oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], c["del"], table, p1, p2, fl)
    if nthreads == 1:
        work(0)
    else:
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(work, range(nthreads)))
        g0 = grids[0]

        def red(name):
            a = getattr(g0, name)
            for t in range(1, nthreads):
                np.add(a, getattr(grids[t], name), out=a)
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(red, ["bias_3d", "rtms_3d", "xcov_3d", "nobs_3d", "bias_2d", "rtms_2d", "xcov_2d", "nobs_2d"]))
    oracle.norm(grids[0], cfg.nlevs, True, 0)


def prepare_cpu_files(files, nr=2):
    out = []
    for c in files:
        n = len(c["lat"])
        d = np.zeros((n, nr), order="F")
        d[:, 0] = c["omf"]                          # gritas.f:562-587 (omf [, oma])
        if nr == 2:
            d[:, 1] = c["oma"]
        cc = dict(c)
        cc["del"] = d
        cc["time"] = np.zeros(n, np.int32)
        out.append(cc)
    return out


def time_cpu(cfg, files, nthreads, steps, warmup):
    from gritas_b200 import synth
    from oracle import oracle
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    grids = [oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, cfg.nr) for _ in range(nthreads)]
    cf = prepare_cpu_files(files, cfg.nr)
    nobs = sum(len(c["lat"]) for c in cf)
    for _ in range(warmup):
        cpu_reference_step(cfg, cf, table, l2d, l3d, p1, p2, nthreads, grids)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(cfg, cf, table, l2d, l3d, p1, p2, nthreads, grids)
    dt = (time.perf_counter() - t0) / steps
    return nobs / dt, dt, nobs


def run_reference(args):
    """--impl reference: CPU restatement of the reference loops on the host cores (kind = "port":
    gritas.x itself needs a Fortran compiler and un-vendored GMAO_Shared, neither is in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gritas_b200 import synth
    cfg = synth.get_config(args.config, args.scale)
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    per_thread = 8 * cfg.im * cfg.jm * cfg.km * len(cfg.kt_list) * (4 * cfg.nr + 1) + (1 << 28)
    nthreads = int(max(1, min(cores, 16, (avail // 2) // per_thread)))
    # bounded sample: about 4 s of CPU work per step (0.33 s per 1.21 M-obs file and thread here)
    nfiles = args.files or int(min(cfg.nfiles, max(nthreads, (4.0 / 0.35) * nthreads * (1.21e6 / cfg.nobs_per_file))))
    nfiles = max(nthreads, (nfiles // nthreads) * nthreads)
    files = gen_files(cfg, 0, nfiles)
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    v, dt, nobs = time_cpu(cfg, files, nthreads, steps, warmup)
    sample = (f"{nfiles} of {cfg.nfiles} files ({nobs} obs) per step, sharded over {nthreads} threads with private "
              f"grids, partial grids summed, then Norm; pre-sort (gritas.f:499-523) not included")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, cfg.nfiles), "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from gritas_b200 import cabi, synth
    from gritas_b200.m_gritas_grids import Grids, nccl_unique_id

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: gritas_b200 has no CPU fallback")
    cabi.load(build_if_missing=True)

    cfg = synth.get_config(args.config, args.scale)
    nfiles = args.files or cfg.nfiles
    files = gen_files(cfg, rank * cfg.nfiles, nfiles)        # before CUDA work: numpy threads only
    nobs_rank = sum(len(c["lat"]) for c in files)
    table0, _, _ = synth.kxkt_table(cfg)
    # observations that must end up in the count grids: QC pass (passive kept, ioptn > 0) and (kx,kt) in the table
    expected_rank = int(sum(np.count_nonzero(np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)) &
                                             (table0[c["kx"] - 1, c["kt"] - 1] != 0)) for c in files))

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    table, l2d, l3d = synth.kxkt_table(cfg)
    mode = (cabi.MODE_FAST if args.mode == "fast" else cabi.MODE_DETERMINISTIC) | cabi.FLAG_ASYNC | cabi.FLAG_HOLD_INPUTS
    if args.no_staging:
        mode |= cabi.FLAG_NO_STAGING
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode, device=local)
    p1, p2 = G.GrITAS_Pint()
    NR = cfg.nr
    IOPT = cabi.IOPT_OMF_OMA if NR == 2 else cabi.IOPT_OMF
    G._ensure(NR, table, p1, p2, l2d, l3d)
    if world > 1:
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        G.comm_init(world, rank, ids[0])
    L = cabi.load()
    stream = torch.cuda.ExternalStream(G.stream(), device=torch.device("cuda", local))

    dev_files = [{k: torch.from_numpy(v).cuda() for k, v in c.items()} for c in files]
    torch.cuda.synchronize()
    # host copies: pinned for the end-to-end arm; the pageable originals are only kept for the CPU baseline sample
    pin_files = None
    if not args.no_e2e:
        pin_files = [{k: torch.from_numpy(v).pin_memory() for k, v in c.items()} for c in files]
    keep = 32 if (rank == 0 and world == 1 and not args.no_cpu_baseline) else 0
    files = files[:keep]
    import ctypes

    def calls_of(cols):
        # argument tuples of gritas_b200_accum_ods, built once: the timed loop is then what a compiled
        # driver's loop over synoptic times would be (one C call per file)
        out = []
        for c in cols:
            p = lambda k: cabi.ptr(c[k])
            out.append((G.handle, int(c["lat"].shape[0]), p("lat"), p("lon"), p("lev"), None, p("kt"), p("kx"),
                        p("qcexcl"), p("omf"), p("oma") if NR == 2 else None, None, None, IOPT, 0, IOPT, 0, -1,
                        -999.0, 999.0, synth.X_PASSIVE, synth.X_PRE_BAD, None, None))
        return out
    dptr = [C_void() for _ in range(10)]
    host_enqueue = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    outs10 = (ctypes.c_void_p * 10)()

    def step(cols, host_out=None, marks=None):
        G.reset()
        if marks is not None:
            marks[0].record(stream)
        th = time.perf_counter()
        for a in cols:
            rc = L.gritas_b200_accum_ods(*a)
            if rc:
                raise RuntimeError(f"accum_ods rc={rc}: {G._err()}")
        host_enqueue.append(time.perf_counter() - th)
        if marks is not None:
            G.barrier()                # the library stream waits for the worker streams (no host sync)
            marks[1].record(stream)
        if world == 1:
            # GrITAS_Norm: on the tiled path K2f accumulates the parked tile lists and finalizes in one kernel
            if host_out is None:
                rc = L.gritas_b200_norm_device(G.handle, 1, -1, *[None] * 5, *dptr[5:])
                if rc:
                    raise RuntimeError(f"norm_device rc={rc}")
            else:
                G.norm_into(host_out)
            if marks is not None:
                marks[2].record(stream)
        else:
            # one collective: tile flush + ncclReduce to rank 0 + finalize, pipelined chunk by chunk
            if marks is not None:
                marks[2].record(stream)
            if host_out is None:
                rc = L.gritas_b200_norm_reduce_device(G.handle, 0, 1, -1, ctypes.cast(outs10, ctypes.POINTER(ctypes.c_void_p)))
                if rc:
                    raise RuntimeError(f"norm_reduce_device rc={rc}: {G._err()}")
            else:
                G.norm_reduce_into(host_out, root=0)

    # ---- device-resident arm ---------------------------------------------------------------
    dev_calls = calls_of(dev_files)
    check = None
    if world > 1:
        # sanity of the multi-rank path: the reduced count grid on rank 0 must hold every rank's accepted obs
        G.reset()
        for a in dev_calls:
            L.gritas_b200_accum_ods(*a)
        raw = G.alloc_grids()
        G.get_raw(None, None, None, None, None, None, None, raw["nobs_3d"])
        mine = torch.tensor([float(raw["nobs_3d"].sum())], dtype=torch.float64, device="cuda")
        dist.all_reduce(mine)
        chk = G.alloc_grids()
        G.reset()
        for a in dev_calls:
            L.gritas_b200_accum_ods(*a)
        G.norm_reduce_into(chk, root=0)
        if rank == 0:
            check = bool(chk["nobs_3d"].sum() == float(mine[0]))
            if not check:
                raise SystemExit(f"multi-rank check failed: reduced count {chk['nobs_3d'].sum()} != {float(mine[0])}")
        del raw, chk
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=20)
    for _ in range(max(3, args.warmup)):
        step(dev_calls)
    c0 = G.counters()
    barrier()
    e0, e1 = ev(), ev()
    marks = [(ev(), ev(), ev()) for _ in range(args.steps)]
    sampler.active = True
    e0.record(stream)
    for s in range(args.steps):
        step(dev_calls, marks=marks[s])
    e1.record(stream)
    barrier()
    sampler.active = False
    c1 = G.counters()
    if world == 1:
        # every accepted observation of the month is in exactly one box (count grid vs the host-side filter)
        raw = G.alloc_grids()
        G.get_raw(None, None, None, None, None, None, None, raw["nobs_3d"])
        check = bool(int(raw["nobs_3d"].sum()) == expected_rank)
        if not check:
            raise SystemExit(f"count check failed: {int(raw['nobs_3d'].sum())} binned, {expected_rank} expected")
        del raw
    host_ms = 1e3 * float(np.mean(host_enqueue[-args.steps:]))
    ms = e0.elapsed_time(e1) / args.steps
    accum_ms = float(np.mean([m[0].elapsed_time(m[1]) for m in marks]))
    flush_ms = float(np.mean([m[1].elapsed_time(m[2]) for m in marks]))
    t = torch.tensor([ms, accum_ms, flush_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, accum_ms, flush_ms = float(t[0]), float(t[1]), float(t[2])
    launches_per_step = (c1["launches"] - c0["launches"]) // args.steps
    value = nobs_rank * world / (ms * 1e-3)

    # ---- end-to-end arm: pinned host columns in, finalized grids out to pinned host arrays ----
    e2e = None
    if not args.no_e2e:
        host_out = G.alloc_grids(pinned=True)
        pin_calls = calls_of(pin_files)
        step(pin_calls, host_out)
        b0 = G.counters()
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step(pin_calls, host_out)
        G.sync()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        sampler.active = False
        b1 = G.counters()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        e2e = {"value": nobs_rank * world / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "h2d_bytes_per_step": (b1["h2d"] - b0["h2d"]) // args.e2e_steps,
               "d2h_bytes_per_step": (b1["d2h"] - b0["d2h"]) // args.e2e_steps,
               "note": "per rank bytes; pinned host columns -> C ABI -> pinned host mean/rms/stdv/xcov/nobs grids"}
        del pin_files
    clocks = sampler.result()

    # ---- roofline of the accumulate path (partition kernels of every file + the tile accumulate) ---------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nbox = cfg.im * cfg.jm * cfg.km * l3d + cfg.im * cfg.jm * l2d
    B_OBS = b_obs(NR)
    grid_bytes = 8 * (1 + 2 * NR) * nbox                     # the sums, once per pass (SURVEY 8d B_grid)
    alg_bytes_per_launch = (nobs_rank * B_OBS + grid_bytes) / nfiles
    staged = L.gritas_b200_tiled(G.handle) == 1
    # the algorithmic bytes hold the grid once (B_grid), so the duration holds the pass over the grid too: at N=1 the
    # whole Norm region (tiled path: K2f, the tile accumulate fused with the finalize; direct / deterministic path: K3)
    path_ms = accum_ms + flush_ms
    launch_ms = path_ms / nfiles
    achieved = alg_bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch from the committed `ncu --set full` capture of this workload (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1e_traffic_cfg2.json")))
        if staged and cfg.name == "cfg2" and nfiles == cfg.nfiles and args.scale == 1.0:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "kernel": (("k_list_obs (K1) per file + k_tile_final<finalize> (K2f: tile accumulate fused with finalize) per Norm" if world == 1 else
                            "k_list_obs (K1) per file + k_tile_final<export> (K2x) per Norm") if staged else
                           ("k_accum_fast" if args.mode == "fast" else "k_box_keys + radix sort + k_segment_sum")),
                "peak_source": peak_src, "avg_launch_ms": launch_ms, "alg_bytes_per_launch": alg_bytes_per_launch,
                "note": f"one launch = the accumulate work of one file: {B_OBS} B/obs x obs per file + (8*(1+2*nr)*boxes)/files; "
                        f"duration = (accumulate region + {'Norm region (K2f: tile accumulate + finalize)' if world == 1 else 'tile flush'}, "
                        f"CUDA events on the library stream) / {nfiles} files"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, nfiles), "mode": args.mode,
                       "per_rank": "each rank bins its own month; N=1: Norm = K2f on the parked tile lists; N>1: K2x (tile sums as compact planes) + ncclReduce to rank 0 + K3x (finalize) pipelined in 8 chunks (rank 0 writes the grids, as the reference has one writer)",
                       "l2": f"inputs ({nobs_rank * b_obs(NR) / 1e9:.1f} GB/step) exceed the 126 MB L2; no flush needed"},
            "accumulate_ms_per_step": accum_ms, "norm_ms_per_step" if world == 1 else "tile_flush_ms_per_step": flush_ms,
            "host_enqueue_ms_per_step": host_ms, "count_check": check, "accepted_obs_per_rank": expected_rank, "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step), "clocks": clocks, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): single thread, like gritas.x ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = len(files)
        v, dt, nobs = time_cpu(cfg, files[:ns], 1, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{ns} of {nfiles} files ({nobs} obs): QC filter + GrITAS_Accum per file, "
                                          f"then GrITAS_Norm, single thread (the reference is serial); its 8-key "
                                          f"pre-sort (gritas.f:499-523) is not included",
                                "host_cores": os.cpu_count()}
    if rank == 0:
        print(json.dumps(line), flush=True)
    G.close()
    if world > 1:
        dist.destroy_process_group()


def C_void():
    import ctypes
    return ctypes.pointer(ctypes.c_void_p())


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
