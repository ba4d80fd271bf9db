#!/usr/bin/env python
"""bench.py -- observations binned per second on the BASELINE.json workload.

A "step" is one pass of the hot path over one month of synthetic conventional ODS columns
(BASELINE.json configs[1]: 124 six-hourly files x ~1.21 M obs, omf+oma in one pass, 1x1 deg x 26
levels, 16 (kt,kx) grids): zero the accumulators, accumulate every file (QC filter + box index +
count/sum/sum-of-squares, fused kernel), finalize mean / rms / stdv.  With N ranks the SAME month is
sharded f mod N (strong scaling, BASELINE.json configs[1]); every rank then finalizes 1/N of the rows,
pulling the parked observations of those rows from all ranks over NVLink peer memory (sharded Norm).
An extra weak-scaling leg (every rank its own month) is reported under "weak".

  value  : whole-job obs/s with the columns already resident in HBM (CUDA events on the library stream)
  e2e    : same step through the C ABI with pinned HOST columns (H2D inside) and the finalized
           grids copied back to pinned host arrays (D2H inside)
  --impl reference : the reference's own loops on the host cores, file-sharded over the threads: oracle/_ref (its Fortran
           routines translated to C statement by statement and compiled here; the reference itself cannot be built:
           Fortran + un-vendored GMAO_Shared), or the hand-written port in oracle/ where that library is missing.
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
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the extra weak-scaling leg")
    ap.add_argument("--norm", default="sharded", choices=["sharded", "reduce"],
                    help="N > 1: sharded Norm over peer memory (default) or the NCCL reduce pipeline")
    return ap.parse_args()


def gen_files(cfg, first, count, workers=8):
    from gritas_b200 import synth

    def one(i):
        c = synth.make_file(cfg, first + i)
        return {k: np.ascontiguousarray(c[k]) for k in COLS_F64 + COLS_I32}
    with ThreadPoolExecutor(max_workers=max(1, min(workers, os.cpu_count() or 1))) as ex:
        return list(ex.map(one, range(count)))


def gen_files_idx(cfg, indices, workers=8):
    from gritas_b200 import synth

    def one(i):
        c = synth.make_file(cfg, i)
        return {k: np.ascontiguousarray(c[k]) for k in COLS_F64 + COLS_I32}
    with ThreadPoolExecutor(max_workers=max(1, min(workers, os.cpu_count() or 1))) as ex:
        return list(ex.map(one, indices))


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
def cpu_reference_step(cfg, files, table, l2d, l3d, p1, p2, nthreads, grids, use_ref=False):
    """One CPU step: the files sharded over `nthreads` workers (private grids each, like independent
    gritas.x processes), QC filter + GrITAS_Accum per file, partial grids summed, GrITAS_Norm."""
    from gritas_b200 import synth
    from oracle import oracle, ref

    def work(t):
        g = grids[t]
        for a in (g.bias_3d, g.rtms_3d, g.xcov_3d, g.nobs_3d, g.bias_2d, g.rtms_2d, g.xcov_2d, g.nobs_2d):
            a[...] = 0.0
        for c in files[t::nthreads]:
            if use_ref:    # the reference's own filter loop and GrITAS_Accum_ (translated to C, oracle/_ref)
                fl = ref.filter_flags(c["qcexcl"], c["time"], c["lon"], ioptn=1, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
                ref.accum(g, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], c["del"], table, p1, p2, fl)
                continue
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
    if use_ref:
        ref.norm(grids[0], True, None)       # GrITAS_Norm_ without the optional threshold, as gritas.f:885-887 calls it
    else:
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


def time_cpu(cfg, files, nthreads, steps, warmup, use_ref=False):
    from gritas_b200 import synth
    from oracle import oracle, ref
    table, l2d, l3d = synth.kxkt_table(cfg)
    p1, p2 = oracle.pint(cfg.plevs)
    if use_ref:
        ref.init(cfg.im, cfg.jm, cfg.km, cfg.nlevs, cfg.plevs, table.shape[0], table.shape[1])
    grids = [oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, cfg.nr) for _ in range(nthreads)]
    cf = prepare_cpu_files(files, cfg.nr)
    nobs = sum(len(c["lat"]) for c in cf)
    for _ in range(warmup):
        cpu_reference_step(cfg, cf, table, l2d, l3d, p1, p2, nthreads, grids, use_ref)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(cfg, cf, table, l2d, l3d, p1, p2, nthreads, grids, use_ref)
    dt = (time.perf_counter() - t0) / steps
    return nobs / dt, dt, nobs


def run_reference(args):
    """--impl reference: the reference's own routines on the host cores -- oracle/_ref, its Fortran translated to C statement
    by statement and compiled here (kind = "reference"); where that library is missing, the hand-written port (kind = "port").
    gritas.x itself needs a Fortran compiler and un-vendored GMAO_Shared, neither is in this image."""
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
    from oracle import ref
    use_ref = ref.available()      # oracle/_ref: built where /root/reference is, travels to the GPU box as a built library
    v, dt, nobs = time_cpu(cfg, files, nthreads, steps, warmup, use_ref)
    what = ("the reference's own filter loop (gritas.f:712-735), GrITAS_Accum_ and GrITAS_Norm_ (m_gritas_grids.f:307-478, 490-656), "
            "translated to C statement by statement (oracle/f2c_subset.py) and compiled with gcc -O2" if use_ref else
            "the hand-written CPU port of the reference loops (oracle/gritas_oracle.c; pinned bit for bit against the translated reference)")
    sample = (f"{nfiles} of {cfg.nfiles} files ({nobs} obs) per step, sharded over {nthreads} threads with private "
              f"grids, partial grids summed, then Norm; pre-sort (gritas.f:499-523) not included; {what}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, cfg.nfiles), "sample": sample},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthreads, "kind": "reference" if use_ref else "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------
def stats3(xs):
    xs = sorted(float(x) for x in xs)
    return {"min": xs[0], "median": xs[len(xs) // 2], "max": xs[-1]}


def run_b200(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from gritas_b200 import cabi, shard, synth
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
    # BASELINE.json configs[1]: ONE month, its files sharded over the ranks f mod N (strong scaling)
    mine = shard.files_for_rank(nfiles, rank, world)
    files = gen_files_idx(cfg, mine)                          # before CUDA work: numpy threads only
    table0, _, _ = synth.kxkt_table(cfg)

    def accepted(fs):
        # observations that must end up in the count grids: QC pass (passive kept, ioptn > 0) and (kx,kt) in the table
        return int(sum(np.count_nonzero(np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)) &
                                        (table0[c["kx"] - 1, c["kt"] - 1] != 0)) for c in fs))
    nobs_rank = sum(len(c["lat"]) for c in files)
    expected_rank = accepted(files)

    def expected_sums(fs):
        # what the finalized grids must add up to: sum x, sum x^2 (per residual column) and sum x1*x2 over the accepted
        # observations, and the sums of their absolute values (the scale of the comparison)
        out = np.zeros(10)
        for c in fs:
            m = np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)) & (table0[c["kx"] - 1, c["kt"] - 1] != 0)
            x1 = c["omf"][m]
            x2 = c["oma"][m] if cfg.nr == 2 else x1
            out += [x1.sum(), x2.sum(), (x1 * x1).sum(), (x2 * x2).sum(), (x1 * x2).sum(),
                    np.abs(x1).sum(), np.abs(x2).sum(), (x1 * x1).sum(), (x2 * x2).sum(), np.abs(x1 * x2).sum()]
        return out
    sums_rank = expected_sums(files)

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    table, l2d, l3d = synth.kxkt_table(cfg)
    mode = (cabi.MODE_FAST if args.mode == "fast" else cabi.MODE_DETERMINISTIC) | cabi.FLAG_ASYNC | cabi.FLAG_HOLD_INPUTS
    if args.no_staging:
        mode |= cabi.FLAG_NO_STAGING
    if world > 1 and args.mode == "fast" and not args.no_staging:
        mode |= cabi.FLAG_FORCE_TILED        # the ranks must agree on the path; a short shard would not trigger the locality probe
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode, device=local)
    p1, p2 = G.GrITAS_Pint()
    NR = cfg.nr
    IOPT = cabi.IOPT_OMF_OMA if NR == 2 else cabi.IOPT_OMF
    G._ensure(NR, table, p1, p2, l2d, l3d)
    norm_kind = "k2f"
    if world > 1:
        ids = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        G.comm_init(world, rank, ids[0])
        blobs = [None] * world
        dist.all_gather_object(blobs, G.peer_export())
        ok = torch.ones(1, device="cuda")
        try:
            if args.norm == "reduce":
                raise RuntimeError("--norm reduce")
            G.peer_attach(world, rank, blobs)
            norm_kind = "sharded-peer"
        except RuntimeError as e:            # no CUDA IPC between the ranks: the NCCL pipeline still works
            ok[0] = 0
            if rank == 0 and args.norm != "reduce":
                print(f"peer_attach failed ({e}); falling back to norm_reduce", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok[0]) == 0:
            norm_kind = "nccl-reduce"
    L = cabi.load()
    stream = torch.cuda.ExternalStream(G.stream(), device=torch.device("cuda", local))
    tot = torch.tensor([float(nobs_rank), float(expected_rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    nobs_total, expected_total = int(tot[0]), int(tot[1])

    def to_dev(fs):
        return [{k: torch.from_numpy(v).cuda() for k, v in c.items()} for c in fs]

    def to_pin(fs):
        return [{k: torch.from_numpy(v).pin_memory() for k, v in c.items()} for c in fs]

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
    det_order = args.mode == "deterministic" and world > 1
    dptr = [C_void() for _ in range(10)]
    outs10 = (ctypes.c_void_p * 10)()
    host_enqueue = []

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def norm(host_out):
        if world == 1:
            # GrITAS_Norm: on the tiled path K2f accumulates the parked tile lists and finalizes in one kernel
            if host_out is None:
                rc = L.gritas_b200_norm_device(G.handle, 1, -1, *[None] * 5, *dptr[5:])
                if rc:
                    raise RuntimeError(f"norm_device rc={rc}: {G._err()}")
            else:
                G.norm_into(host_out)
        elif norm_kind == "sharded-peer":
            # every rank finalizes its slab of rows, pulling the parked entries of those rows from all ranks over NVLink
            if host_out is None:
                rc = L.gritas_b200_norm_sharded_device(G.handle, 1, -1, ctypes.cast(outs10, ctypes.POINTER(ctypes.c_void_p)))
                if rc:
                    raise RuntimeError(f"norm_sharded_device rc={rc}: {G._err()}")
            else:
                G.norm_sharded_into(host_out)
        else:
            if host_out is None:
                rc = L.gritas_b200_norm_reduce_device(G.handle, 0, 1, -1, ctypes.cast(outs10, ctypes.POINTER(ctypes.c_void_p)))
                if rc:
                    raise RuntimeError(f"norm_reduce_device rc={rc}: {G._err()}")
            else:
                G.norm_reduce_into(host_out, root=0)

    def step(cols, host_out=None, marks=None):
        G.reset()
        if marks is not None:
            marks[0].record(stream)
        th = time.perf_counter()
        for ci, a in enumerate(cols):
            if det_order:      # deterministic mode on several ranks: every call carries its global file index
                L.gritas_b200_set_call_order(G.handle, mine[ci] if ci < len(mine) else ci)
            rc = L.gritas_b200_accum_ods(*a)
            if rc:
                raise RuntimeError(f"accum_ods rc={rc}: {G._err()}")
        host_enqueue.append(time.perf_counter() - th)
        if marks is not None:
            G.barrier()                # the library stream waits for the worker streams (no host sync)
            marks[1].record(stream)
        norm(host_out)
        if marks is not None:
            marks[2].record(stream)

    def slab_sums(host):
        """[accepted observations, n*mean (x1, x2), n*rms^2 (x1, x2), n*<x1 x2>] summed over this rank's slab of the finalized
        3-D grids (the whole grids at N=1): the statistics of every box, folded back into the sums they came from"""
        if world == 1 or norm_kind != "sharded-peer":
            if not (world == 1 or rank == 0):
                return np.zeros(6)
            pieces = [(slice(None), slice(None))]
        else:
            a3, b3, _, _ = G.slab(rank, world)
            pieces = []
            for l in range(a3 // cfg.jm, (b3 - 1) // cfg.jm + 1 if b3 > a3 else 0):
                ja, jb = max(a3, l * cfg.jm) - l * cfg.jm, min(b3, (l + 1) * cfg.jm) - l * cfg.jm
                pieces.append((slice(ja, jb), l))
        tot = np.zeros(6)
        for js, ls in pieces:
            n = host["nobs_3d"][:, js, :, ls]
            w = np.where(n > 0, n, 0.0)
            i2 = NR - 1
            b1, b2 = host["bias_3d"][:, js, :, ls, 0], host["bias_3d"][:, js, :, ls, i2]
            r1, r2 = host["rtms_3d"][:, js, :, ls, 0], host["rtms_3d"][:, js, :, ls, i2]
            xy = host["xcov_3d"][:, js, :, ls, 0] if NR == 2 else r1 * r1
            z = lambda a: np.where(n > 0, a, 0.0)
            tot += [n.sum(), (w * z(b1)).sum(), (w * z(b2)).sum(), (w * z(r1) ** 2).sum(), (w * z(r2) ** 2).sum(), (w * z(xy)).sum()]
        return tot

    def timed_device(calls, steps, warmup):
        for _ in range(warmup):
            step(calls)
        c0 = G.counters()
        barrier()
        e0, e1 = ev(), ev()
        marks = [(ev(), ev(), ev()) for _ in range(steps)]
        sampler.active = True
        e0.record(stream)
        n_enq = len(host_enqueue)
        for s in range(steps):
            step(calls, marks=marks[s])
        e1.record(stream)
        enq = list(host_enqueue[n_enq:])      # host time of the Accum calls of the timed device-resident steps only
        barrier()
        sampler.active = False
        c1 = G.counters()
        acc = [m[0].elapsed_time(m[1]) for m in marks]
        nrm = [m[1].elapsed_time(m[2]) for m in marks]
        t = torch.tensor([e0.elapsed_time(e1) / steps, float(np.mean(acc)), float(np.mean(nrm))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return {"ms": float(t[0]), "accum_ms": float(t[1]), "norm_ms": float(t[2]), "host_enqueue_ms": 1e3 * float(np.mean(enq)), "accum_spread": stats3(acc),
                "norm_spread": stats3(nrm), "launches_per_step": (c1["launches"] - c0["launches"]) // steps}

    def timed_e2e(calls, host_out, steps):
        step(calls, host_out)
        b0 = G.counters()
        barrier()
        sampler.active = True
        t0 = time.perf_counter()
        for _ in range(steps):
            step(calls, host_out)
        G.sync()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        sampler.active = False
        b1 = G.counters()
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        by = torch.tensor([float(b1["h2d"] - b0["h2d"]), float(b1["d2h"] - b0["d2h"])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(by)
        return float(tt[0]), int(by[0]) // steps, int(by[1]) // steps

    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=20)

    # ---- device-resident arm ---------------------------------------------------------------
    dev_files = to_dev(files)
    torch.cuda.synchronize()
    dev_calls = calls_of(dev_files)
    dev = timed_device(dev_calls, args.steps, max(3, args.warmup))
    ms, accum_ms, norm_ms = dev["ms"], dev["accum_ms"], dev["norm_ms"]
    value = nobs_total / (ms * 1e-3)

    # every accepted observation of the month is in exactly one box (count grid vs the host-side filter)
    host_out = G.alloc_grids(pinned=True)
    step(dev_calls, host_out)
    cnt = torch.tensor(np.concatenate([slab_sums(host_out), sums_rank]), dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    cnt = cnt.cpu().numpy()
    check = bool(int(cnt[0]) == expected_total)
    if not check:
        raise SystemExit(f"count check failed: {int(cnt[0])} binned, {expected_total} expected")
    # ... and the means / mean squares / mean products of all boxes add up to the month's sums (every rank's slab, any N)
    sum_err = float(np.max(np.abs(cnt[1:6] - cnt[6:11]) / np.maximum(cnt[11:16], 1e-300)))
    if not sum_err < 1e-9:
        raise SystemExit(f"sum check failed: finalized grids do not add up to the month's sums (relative error {sum_err:.3e})")

    # ---- end-to-end arm: pinned host columns in, finalized grids out to pinned host arrays ----
    e2e = None
    if not args.no_e2e:
        pin_files = to_pin(files)
        dt, h2d, d2h = timed_e2e(calls_of(pin_files), host_out, args.e2e_steps)
        e2e = {"value": nobs_total / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "note": "bytes summed over the ranks; pinned host columns -> gritas_b200_accum_ods -> Norm -> pinned host "
                       "mean/rms/stdv/xcov/nobs grids" + ("" if world == 1 else " (every rank copies its slab of rows)")}
        try:   # the box's host <-> device copy ceiling with this many ranks copying at once (tools/h2d_probe.py, committed runs)
            pr = json.load(open(os.path.join(ROOT, "profiles", f"r2_h2d_probe_n{world}.json")))
            floor = h2d / (pr["h2d_all_gbs_total"] * 1e9) + d2h / (pr["d2h_all_gbs_total"] * 1e9)
            e2e["copy_floor_ms"] = floor * 1e3
            e2e["frac_of_copy_floor"] = floor / dt
            e2e["copy_floor_source"] = f"profiles/r2_h2d_probe_n{world}.json: {pr['h2d_all_gbs_total']:.0f} GB/s H2D, {pr['d2h_all_gbs_total']:.0f} GB/s D2H in total with {world} rank(s) copying"
        except Exception:
            pass
        del pin_files

    # ---- the drop-in's end to end: what an untouched gritas.f gets through fortran/m_gritas_b200.F90 ----
    # pageable columns, del built and ob_flag filtered on the host as gritas.f:562-587, 710-739 do, GrITAS_Accum
    # (ob_flag inout, synchronous), GrITAS_Norm into pageable arrays
    e2e_pageable = None
    if not args.no_e2e and world == 1:
        pfiles = []
        for c in files:
            d = np.empty((len(c["lat"]), NR), order="F")
            d[:, 0] = c["omf"]
            if NR == 2:
                d[:, 1] = c["oma"]
            # gritas.f:710-739 as the untouched driver runs it on the host (no time / longitude window, ioptn > 0)
            fl = np.where(np.isin(c["qcexcl"], (0, synth.X_PASSIVE, synth.X_PRE_BAD)), 0, 1).astype(np.int32)
            pfiles.append((c, d, fl))

        def pageable_leg(mode_flags):
            Gp = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_FAST | mode_flags, device=local)
            Gp._ensure(NR, table, p1, p2, l2d, l3d)
            bad = ctypes.c_double(0.0)
            pcalls = [(Gp.handle, len(fl), cabi.ptr(c["lat"]), cabi.ptr(c["lon"]), cabi.ptr(c["lev"]), cabi.ptr(d), cabi.ptr(c["kx"]),
                       cabi.ptr(c["kt"]), cabi.ptr(fl), ctypes.byref(bad)) for c, d, fl in pfiles]
            pout = Gp.alloc_grids()

            def pstep():
                Gp.reset()
                for a in pcalls:
                    rc = L.gritas_b200_accum(*a)      # ob_flag is inout: the flags it sets (not-in-table) only re-reject the same obs
                    if rc:
                        raise RuntimeError(f"accum rc={rc}: {Gp._err()}")
                Gp.norm_into(pout)
            pstep()
            c0 = Gp.counters()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                pstep()
            dtp = (time.perf_counter() - t0) / args.e2e_steps
            c1 = Gp.counters()
            ok = bool(int(pout["nobs_3d"].sum()) == expected_total)
            Gp.close()
            return dtp, (c1["h2d"] - c0["h2d"]) // args.e2e_steps, (c1["d2h"] - c0["d2h"]) // args.e2e_steps, ok
        dtp, ph2d, pd2h, ok = pageable_leg(0)
        dta, _, _, oka = pageable_leg(cabi.FLAG_ASYNC | cabi.FLAG_NO_WRITEBACK)
        e2e_pageable = {"value": nobs_total / dtp, "unit": UNIT, "ms_per_step": dtp * 1e3, "steps": args.e2e_steps,
                        "h2d_bytes_per_step": ph2d, "d2h_bytes_per_step": pd2h, "count_check": ok and oka,
                        "async_no_writeback": {"value": nobs_total / dta, "ms_per_step": dta * 1e3,
                                               "note": "B200_SetFlags(async=.true., writeback=.false.): a driver without -lb"},
                        "note": "pageable numpy columns + host-built del and ob_flag -> gritas_b200_accum (ob_flag inout, "
                                "synchronous) per file -> gritas_b200_norm into pageable arrays: the call sequence of "
                                "fortran/m_gritas_b200.F90 under an untouched gritas.f; staging through the library's host copy pool"}
        del pfiles

    # ---- weak leg (extra): every rank bins a whole month of its own, one sharded Norm over all of them ----
    weak = None
    if world > 1 and not args.no_weak:
        del dev_files, dev_calls
        wfiles = gen_files_idx(cfg, [rank * cfg.nfiles + i for i in range(nfiles)])
        wn = torch.tensor([float(sum(len(c["lat"]) for c in wfiles))], dtype=torch.float64, device="cuda")
        dist.all_reduce(wn)
        wdev = to_dev(wfiles)
        del wfiles
        w = timed_device(calls_of(wdev), max(3, args.steps // 2), 2)
        weak = {"value": float(wn[0]) / (w["ms"] * 1e-3), "unit": UNIT, "ms_per_step": w["ms"], "accumulate_ms_per_step": w["accum_ms"],
                "norm_ms_per_step": w["norm_ms"], "note": "weak scaling: every rank bins its own month (N months per step), one sharded Norm"}
        del wdev
    clocks = sampler.result()

    # ---- roofline of the path: K1 of every file + the Norm kernel that accumulates the parked tiles and finalizes ----
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nbox = cfg.im * cfg.jm * cfg.km * l3d + cfg.im * cfg.jm * l2d
    B_OBS = b_obs(NR)
    grid_bytes = 8 * (1 + 2 * NR) * nbox                     # the sums, once per pass (SURVEY 8d B_grid)
    # per GPU and launch: this rank's observations + its share of the grid pass (each rank finalizes 1/N of the rows)
    nf_rank = max(1, len(files))
    alg_bytes_per_launch = (nobs_total / world * B_OBS + grid_bytes / world) / nf_rank
    staged = L.gritas_b200_tiled(G.handle) == 1
    # the algorithmic bytes hold the grid once (B_grid), so the duration holds the pass over the grid too: the whole Norm
    # region (tiled path: the tile kernel = accumulate of the parked lists fused with the finalize, at N > 1 including the
    # pull over NVLink and the two barriers; direct / deterministic path: K3)
    path_ms = accum_ms + norm_ms
    launch_ms = path_ms / nf_rank
    achieved = alg_bytes_per_launch / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:   # DRAM bytes per launch from this round's `ncu --set full` capture of this workload (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic_cfg2.json")))
        if staged and world == 1 and cfg.name == "cfg2" and nfiles == cfg.nfiles and args.scale == 1.0:
            traffic = tr["dram_bytes_per_launch"]
    except Exception:
        pass
    kern = ("k_list_obs (K1) per file + k_tile_final (tile accumulate fused with finalize"
            + ("" if world == 1 else ", entries pulled from all ranks over NVLink peer memory") + ") per Norm") if staged else \
           ("k_accum_fast" if args.mode == "fast" else "k_box_keys + radix sort + k_segment_sum")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kern,
                "peak_source": peak_src, "avg_launch_ms": launch_ms, "alg_bytes_per_launch": alg_bytes_per_launch,
                "note": f"per GPU; one launch = the accumulate work of one file: {B_OBS} B/obs x obs per file + this rank's share of "
                        f"(8*(1+2*nr)*boxes) / its files; duration = (accumulate region + the whole Norm region, CUDA events on "
                        f"the library stream, max over ranks) / {nf_rank} files per rank"}

    per_rank = {1: "N=1: Norm = K2f on the parked tile lists",
                }.get(world, f"file f -> rank f mod {world}; Norm = {norm_kind}: " +
                      ("every rank finalizes 1/N of the rows, its tile kernel pulls the parked entries of those rows from all "
                       "ranks' lists over NVLink (CUDA IPC peer mappings); barriers = two one-element NCCL all-reduces"
                       if norm_kind == "sharded-peer" else
                       "K2x (tile sums as compact planes) + ncclReduce to rank 0 + K3x (finalize) pipelined in 8 chunks"))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, nfiles), "mode": args.mode, "sharding": per_rank,
                       "l2": f"inputs ({nobs_rank * b_obs(NR) / 1e9:.2f} GB/step per rank) exceed the 126 MB L2; no flush needed"},
            "accumulate_ms_per_step": accum_ms, "norm_ms_per_step": norm_ms,
            "spread": {"accumulate_ms": dev["accum_spread"], "norm_ms": dev["norm_spread"], "note": "rank 0, per timed step"},
            "host_enqueue_ms_per_step": dev["host_enqueue_ms"], "count_check": check,
            "sum_check": {"ok": True, "max_rel_err": sum_err,
                          "what": "sum over all boxes of n*mean, n*rms^2, n*<x1 x2> of the finalized grids (every rank's slab) against "
                                  "the sums over the accepted observations, relative to the sums of absolute values"},
            "accepted_obs": expected_total, "gpu_launches": int(dev["launches_per_step"] * args.steps),
            "gpu_launches_per_step": int(dev["launches_per_step"]), "clocks": clocks, "roofline": roofline}
    if world == 1:
        line["scaling_note"] = "N=1: one GPU bins the whole month; N>1 lines shard the same month (strong scaling)"
    if e2e:
        line["e2e"] = e2e
    if e2e_pageable:
        line["e2e_pageable"] = e2e_pageable
    if weak:
        line["weak"] = weak

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): single thread, like gritas.x ----
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(32, len(files))
        from oracle import ref
        use_ref = ref.available()          # oracle/_ref: the reference's own routines, translated to C and compiled (else the port)
        v, dt, nobs = time_cpu(cfg, files[:ns], 1, 1, 1, use_ref)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "reference" if use_ref else "port",
                                "sample": f"{ns} of {nfiles} files ({nobs} obs): QC filter + GrITAS_Accum per file, "
                                          f"then GrITAS_Norm, single thread (the reference is serial); its 8-key "
                                          f"pre-sort (gritas.f:499-523) is not included; " +
                                          ("the reference's own gritas.f:712-735 / GrITAS_Accum_ / GrITAS_Norm_ translated to C "
                                           "statement by statement (oracle/f2c_subset.py), gcc -O2" if use_ref else
                                           "hand-written port (oracle/gritas_oracle.c), pinned bit for bit against the translated reference"),
                                "host_cores": os.cpu_count()}
    if rank == 0:
        emit(line)
    G.close()
    if world > 1:
        dist.destroy_process_group()


def C_void():
    import ctypes
    return ctypes.pointer(ctypes.c_void_p())


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    # libraries (NCCL prints its version banner) write to stdout: keep fd 1 for the JSON line only
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
