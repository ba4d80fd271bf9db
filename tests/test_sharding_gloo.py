"""Multi-rank host logic on CPU (gloo, world_size 2): files sharded round-robin, partial
count / sum / sum-of-squares grids summed with one all-reduce, then GrITAS_Norm -- must give the
single-process result (counts bit-exact; sums differ only by fp64 re-association)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gritas_b200 import shard, synth
from oracle import oracle
from tests import helpers as H


def test_files_for_rank_partitions_exactly():
    for nfiles in (0, 1, 7, 124):
        for world in (1, 2, 3, 8):
            parts = [shard.files_for_rank(nfiles, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(nfiles))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for r, p in enumerate(parts):
                assert all(shard.owner_of(i, world) == r for i in p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nfiles, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.get_config("cfg2", 0.004)
    cfg.im, cfg.jm = 48, 25
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 2)
    for fi in shard.files_for_rank(nfiles, rank, world):
        c = synth.make_file(cfg, fi)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), table, p1, p2,
                     H.filter_in(c, 101))
    # the one exchange step of the path: sum the raw accumulators over ranks, then finalize
    for nm in ("bias_3d", "rtms_3d", "xcov_3d", "nobs_3d"):
        t = torch.from_numpy(getattr(g, nm))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    oracle.norm(g, cfg.nlevs, True, 0)
    if rank == 0:
        q.put({nm: getattr(g, nm).copy() for nm in ("bias_3d", "rtms_3d", "stdv_3d", "xcov_3d", "nobs_3d")})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_then_norm_matches_single_process():
    world, nfiles = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nfiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = synth.get_config("cfg2", 0.004)
    cfg.im, cfg.jm = 48, 25
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    assert np.array_equal(got["nobs_3d"], ref.nobs_3d)
    table, l2d, l3d, *_ = H.setup(cfg)
    got.update({k: getattr(ref, k) for k in ("bias_2d", "rtms_2d", "stdv_2d", "xcov_2d", "nobs_2d")})
    H.compare(got, ref, 0, l3d, 2, exact=False, tol=1e-12)


def test_slab_rows_partition_every_row_once():
    for (im, jm, km, l2d, l3d) in ((72, 46, 26, 0, 16), (72, 46, 26, 3, 4), (8, 5, 2, 1, 1), (360, 181, 26, 0, 16), (16, 9, 1, 2, 0)):
        for world in (1, 2, 3, 5, 8):
            rows = [shard.slab_rows(im, jm, km, l2d, l3d, r, world) for r in range(world)]
            assert rows[0][0] == 0 and rows[0][2] == 0
            assert rows[-1][1] == jm * l3d and rows[-1][3] == jm * l2d
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0] and a[3] == b[2]          # contiguous, no gap, no overlap
            for a3, b3, a2, b2 in rows:
                assert a3 <= b3 and a2 <= b2


def _worker_sharded(rank, world, port, nfiles, q):
    """The sharded Norm's host logic on CPU: every rank accumulates its files, the raw sums of the rows a rank owns are
    summed over the ranks (here: all_reduce, the GPU path pulls the parked entries instead), every rank finalizes and
    reports ONLY its slab of rows; the stitched slabs must be the single-process grids."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.get_config("cfg2", 0.004)
    cfg.im, cfg.jm = 48, 25
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    g = oracle.Grids(cfg.im, cfg.jm, cfg.km, l2d, l3d, 2)
    for fi in shard.files_for_rank(nfiles, rank, world):
        c = synth.make_file(cfg, fi)
        oracle.accum(g, cfg.nlevs, c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 2), table, p1, p2,
                     H.filter_in(c, 101))
    for nm in ("bias_3d", "rtms_3d", "xcov_3d", "nobs_3d"):
        dist.all_reduce(torch.from_numpy(getattr(g, nm)), op=dist.ReduceOp.SUM)
    oracle.norm(g, cfg.nlevs, True, 0)
    a3, b3, _, _ = shard.slab_rows(cfg.im, cfg.jm, cfg.km, l2d, l3d, rank, world)
    part = {}
    for nm in ("bias_3d", "rtms_3d", "stdv_3d", "xcov_3d", "nobs_3d"):
        full = getattr(g, nm)
        out = np.full_like(full, np.nan)
        for jl in range(a3, b3):
            j, l = jl % cfg.jm, jl // cfg.jm
            out[:, j, :, l] = full[:, j, :, l]
        part[nm] = out
    q.put((rank, (a3, b3), part))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_norm_slabs_stitch_to_the_single_process_grids():
    world, nfiles = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, nfiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = synth.get_config("cfg2", 0.004)
    cfg.im, cfg.jm = 48, 25
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    table, l2d, l3d, *_ = H.setup(cfg)
    merged = {nm: np.full_like(getattr(ref, nm), np.nan) for nm in parts[0][2]}
    for rank, (a3, b3), part in parts:
        for nm, arr in part.items():
            own = ~np.isnan(arr)
            assert not np.any(own & ~np.isnan(merged[nm])), "a row written by two ranks"
            merged[nm][own] = arr[own]
    for nm in merged:
        assert not np.isnan(merged[nm]).any(), "a row written by no rank"
    assert np.array_equal(merged["nobs_3d"], ref.nobs_3d)
    merged.update({k: getattr(ref, k) for k in ("bias_2d", "rtms_2d", "stdv_2d", "xcov_2d", "nobs_2d")})
    H.compare(merged, ref, 0, l3d, 2, exact=False, tol=1e-12)
