"""Two ranks on two GPUs: files sharded f mod 2, partial accumulators reduced to rank 0 inside the pipelined
multi-rank Norm, result compared with the single-process oracle.  Skipped on boxes with one GPU."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _worker(rank, world, nfiles, uid_q, out_q, mode):
    import torch

    from gritas_b200 import shard, synth
    from gritas_b200.m_gritas_grids import Grids, nccl_unique_id
    torch.cuda.set_device(rank)
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode, device=rank)
    G._ensure(2, table, p1, p2, l2d, l3d)
    if rank == 0:
        uid = nccl_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    G.comm_init(world, rank, uid)
    for fi in shard.files_for_rank(nfiles, rank, world):
        c = synth.make_file(cfg, fi)
        G.accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
    out = G.alloc_grids()
    G.norm_reduce_into(out, root=0)
    again = None
    if mode & 0x1000:
        # tiled path: the multi-rank Norm (K2x -> ncclReduce of compact planes -> K3x) leaves the accumulation state as it
        # was; the record-based pipeline of the other paths reduces in place and is a one-shot operation
        again = G.alloc_grids()
        G.norm_reduce_into(again, root=0)
    if rank == 0:
        for k in (out if again is not None else ()):
            if k.startswith("nobs"):
                assert np.array_equal(out[k], again[k]), k
            else:
                assert np.allclose(out[k], again[k], rtol=1e-9, atol=1e-9, equal_nan=True), k
        out_q.put({k: v.copy() for k, v in out.items()})
    G.close()


@pytest.mark.parametrize("mode", [0, 1, 0x1000])     # fast (direct path at this size), deterministic, fast on the tiled path
def test_two_rank_norm_reduce_matches_oracle(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from gritas_b200 import synth
    world, nfiles = 2, 5
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, nfiles, uid_q, out_q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = out_q.get(timeout=150)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    table, l2d, l3d, *_ = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(got, ref, l2d, l3d, 2, exact=False, tol=1e-12)      # sums re-associated across ranks: not bit-exact


def _worker_sharded(rank, world, nfiles, uid_q, blob_qs, out_q, mode):
    import torch

    from gritas_b200 import shard, synth
    from gritas_b200.m_gritas_grids import Grids, nccl_unique_id
    torch.cuda.set_device(rank)
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode | 0x100, device=rank)     # ASYNC: nothing waits on the host
    G._ensure(2, table, p1, p2, l2d, l3d)
    if rank == 0:
        uid = nccl_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    G.comm_init(world, rank, uid)
    blob = G.peer_export()
    for q in blob_qs:
        q.put((rank, blob))
    got = dict(blob_qs[rank].get(timeout=120) for _ in range(world))
    G.peer_attach(world, rank, [got[r] for r in range(world)])       # CUDA IPC mappings of the peer's lists / records
    outs = []
    for rep in range(2):                                              # twice: the barriers order the reset against the peer's reads
        G.reset()
        for fi in shard.files_for_rank(nfiles, rank, world):
            c = synth.make_file(cfg, fi)
            G.accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        out = G.alloc_grids()
        G.norm_sharded_into(out)
        outs.append(out)
    for k in outs[0]:          # same inputs, same slab; the fast path is order-free, so only the counts repeat bit for bit
        if k.startswith("nobs"):
            assert np.array_equal(outs[0][k], outs[1][k]), k
        else:
            assert np.allclose(outs[0][k], outs[1][k], rtol=1e-9, atol=1e-9, equal_nan=True), k
    out_q.put((rank, G.slab(rank, world), {k: v.copy() for k, v in outs[1].items()}))
    G.close()


# fast on the tiled path (entries pulled over NVLink), the same with the planes exchange forced, deterministic (entries merged
# by sequence number)
@pytest.mark.parametrize("mode,planes", [(0x1000, False), (0x1000, True), (1, False)])
def test_two_rank_sharded_norm_over_ipc_matches_oracle(mode, planes, monkeypatch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    if planes:
        monkeypatch.setenv("GRITAS_B200_PLANES_ABOVE", "0")      # inherited by the spawned ranks

    from gritas_b200 import synth
    world, nfiles = 2, 5
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    blob_qs = [ctx.Queue() for _ in range(world)]
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, nfiles, uid_q, blob_qs, out_q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [out_q.get(timeout=200) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(nfiles)]
    table, l2d, l3d, *_ = H.setup(cfg)
    # stitch the slabs: rank r wrote rows [jl0, jl1) of every 3-D array (jl = j + jm*l)
    merged = {k: np.full_like(v, -777.0) for k, v in parts[0][2].items()}
    for rank, (a3, b3, a2, b2), out in parts:
        for k, v in out.items():
            if not k.endswith("3d"):
                continue
            for jl in range(a3, b3):
                j, l = jl % cfg.jm, jl // cfg.jm
                merged[k][:, j, :, l] = v[:, j, :, l]
    for k in list(merged):
        if k.endswith("2d"):
            merged[k][...] = parts[0][2][k]
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(merged, ref, 0, l3d, 2, exact=False, tol=1e-12 if mode == 1 else 1e-6)
