"""Sharded multi-rank GrITAS_Norm over peer memory, emulated on ONE GPU: N handles in one process stand for N ranks
(files f mod N, gritas.f:389-393 on each share); every "rank" finalizes its slab of rows by pulling the parked
observations of those rows from all handles' tile lists (k_tile_final<TILE_PEERS>).  The union of the slabs must be
the single-process oracle result (gritas.f:885-887 after the whole file loop).  The two-process / two-GPU variant with
CUDA IPC mappings is tests/test_gpu_multirank.py."""
import numpy as np
import pytest

from gritas_b200 import cabi, shard, synth
from gritas_b200.m_gritas_grids import Grids
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _run(cfg, files, nr, world, mode, iopt, nrmlz=True, ntresh=None, env_slots=None):
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    Gs = []
    for r in range(world):
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=mode)
        G._ensure(nr, table, p1, p2, l2d, l3d)
        Gs.append(G)
    blobs = [G.peer_export() for G in Gs]
    for r, G in enumerate(Gs):
        G.peer_attach(world, r, blobs)
    for r, G in enumerate(Gs):
        for fi in shard.files_for_rank(len(files), r, world):
            G.accum_ods(files[fi], iopt=iopt, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
    for G in Gs:
        G.sync_no_flush()
    for G in Gs:
        G.export_planes()        # (no communicator here: phase 1 of the planes exchange, used when the data are dense)
    out = Gs[0].alloc_grids()
    for a in out.values():
        a[...] = -777.0          # every box must be written by exactly one rank
    for G in Gs:
        G.norm_sharded_into(out, nrmlz, ntresh)
    slabs = [G.slab(r, world) for r, G in enumerate(Gs)]
    tiled = [cabi.load().gritas_b200_tiled(G.handle) for G in Gs]
    for G in Gs:
        G.close()
    return out, slabs, tiled, (l2d, l3d)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("nr", [1, 2])
def test_sharded_norm_tiled_matches_oracle(world, nr):
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(9)]
    iopt = 101 if nr == 2 else 1
    out, slabs, tiled, (l2d, l3d) = _run(cfg, files, nr, world, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, iopt)
    assert all(t == 1 for t in tiled)
    # the slabs partition the rows
    assert slabs[0][0] == 0 and slabs[-1][1] == cfg.jm * l3d
    for a, b in zip(slabs, slabs[1:]):
        assert a[1] == b[0] and a[3] == b[2]
    ref, _ = H.oracle_run(cfg, files, nr, [H.filter_in(c, iopt) for c in files])
    H.compare(out, ref, l2d, l3d, nr, exact=False, tol=1e-6)


def test_sharded_norm_surface_and_upper_air():
    """2-D and 3-D grids: the slab boundary may fall into either part."""
    cfg = synth.get_config("cfg1", 0.03)
    cfg.im, cfg.jm = 72, 46
    cfg.kt_list = [44, 4, 5, 1, 2, 3, 44]
    cfg.kx_lists = [[120, 130], [220], [220], [181, 187], [181], [180], [130]]
    files = []
    for i in range(5):
        c = synth.make_file(synth.get_config("cfg1", 0.03), i)
        rng = np.random.default_rng(100 + i)
        row = rng.integers(0, 7, len(c["lat"]))
        c["kt"] = np.asarray(cfg.kt_list, np.int32)[row]
        c["kx"] = np.asarray([120, 220, 220, 181, 181, 180, 130], np.int32)[row]
        files.append(c)
    for world in (2, 5):
        out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 2, world, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, 101)
        assert slabs[-1][3] == cfg.jm * l2d
        ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
        H.compare(out, ref, l2d, l3d, 2, exact=False, tol=1e-6)


@pytest.mark.parametrize("mode", [cabi.MODE_DETERMINISTIC, cabi.MODE_FAST | cabi.FLAG_NO_STAGING])
def test_sharded_norm_record_path(mode):
    """No tile lists (deterministic / direct path): the slab's records are summed over the ranks in rank order."""
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(5)]
    out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 2, 3, mode, 101)
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=1e-12 if H.is_det(mode) else 1e-6)


def test_sharded_norm_mixed_lists_and_records():
    """A rank whose lists overflowed (direct accumulate, dirty tiles) and ranks with everything parked."""
    import os
    cfg = synth.get_config("cfg5", 0.05)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(6)]
    os.environ["GRITAS_B200_STAGE_SLOTS"] = str(1 << 16)       # tiny lists: the clustered tiles overflow
    try:
        out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 2, 3, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, 101)
    finally:
        del os.environ["GRITAS_B200_STAGE_SLOTS"]
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=1e-6)


def test_sharded_norm_leaves_state_and_repeats():
    """Like the one-rank Norm on the tiled path, the sharded Norm does not change the accumulation state."""
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(4)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    Gs = [Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_FAST | cabi.FLAG_FORCE_TILED) for _ in range(2)]
    for G in Gs:
        G._ensure(2, table, p1, p2, l2d, l3d)
    blobs = [G.peer_export() for G in Gs]
    for r, G in enumerate(Gs):
        G.peer_attach(2, r, blobs)
    res = []
    for rep in range(2):
        for r, G in enumerate(Gs):
            for fi in shard.files_for_rank(len(files), r, 2)[rep::2]:
                G.accum_ods(files[fi], iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
        for G in Gs:
            G.sync_no_flush()
        out = Gs[0].alloc_grids()
        for G in Gs:
            G.norm_sharded_into(out)
        res.append(out)
    # first Norm saw files {0, 1}, second all four
    ref2, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(res[1], ref2, l2d, l3d, 2, exact=False, tol=1e-6)
    ref1, _ = H.oracle_run(cfg, files[:2], 2, [H.filter_in(c, 101) for c in files[:2]])
    H.compare(res[0], ref1, l2d, l3d, 2, exact=False, tol=1e-6)
    for G in Gs:
        G.close()


def test_peer_attach_rejects_mismatched_ranks():
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    A = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_FAST | cabi.FLAG_FORCE_TILED)
    B = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_DETERMINISTIC)
    for G in (A, B):
        G._ensure(2, table, p1, p2, l2d, l3d)
    blobs = [A.peer_export(), B.peer_export()]
    for r, G in enumerate((A, B)):
        with pytest.raises(RuntimeError, match="differs from rank 0"):
            G.peer_attach(2, r, blobs)
    A.close()
    B.close()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_norm_planes_exchange(world, monkeypatch):
    """When the ranks park more entries than ~1.5 per box, the sharded Norm exchanges sum planes instead of entries: every rank
    exports its tiles' {n, sums} (K2x) and the owner of a slab sums the ranks' planes through peer memory (K3p).  Forced here
    with GRITAS_B200_PLANES_ABOVE=0; the threshold itself is exercised with dense data (few boxes, many observations)."""
    monkeypatch.setenv("GRITAS_B200_PLANES_ABOVE", "0")
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(9)]
    out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 2, world, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, 101)
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=1e-6)
    monkeypatch.delenv("GRITAS_B200_PLANES_ABOVE")


def test_sharded_norm_dense_data_takes_the_planes_path():
    """8 x 5 x 26 x 16 boxes, ~290 K accepted observations on 3 ranks: ~6 entries per box and rank -> planes."""
    cfg = synth.get_config("cfg2", 0.1)
    cfg.im, cfg.jm = 8, 5
    files = [synth.make_file(cfg, i) for i in range(3)]
    out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 1, 3, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, 1)
    ref, _ = H.oracle_run(cfg, files, 1, [H.filter_in(c, 1) for c in files])
    H.compare(out, ref, l2d, l3d, 1, exact=False, tol=1e-6)


def test_static_slab_matches_the_python_twin():
    """gritas_b200_slab before any sharded Norm = shard.slab_rows (the partition the CPU gloo test uses)."""
    for (im, jm, lists) in ((72, 46, None), (24, 13, None)):
        cfg = synth.get_config("cfg1", 0.01)
        cfg.im, cfg.jm = im, jm
        cfg.kt_list = [44, 4, 5, 1, 2, 3, 44]
        cfg.kx_lists = [[120, 130], [220], [220], [181, 187], [181], [180], [130]]
        table, l2d, l3d, p1, p2 = H.setup(cfg)
        G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=cabi.MODE_FAST | cabi.FLAG_FORCE_TILED)
        G._ensure(1, table, p1, p2, l2d, l3d)
        for world in (1, 2, 3, 8):
            for r in range(world):
                assert G.slab(r, world) == shard.slab_rows(cfg.im, cfg.jm, cfg.km, l2d, l3d, r, world)
        G.close()


def test_sharded_norm_with_an_idle_rank_and_a_direct_path_rank():
    """Rank 2 gets no file at all; rank 1 runs the direct kernel (its files are too small for the locality probe, so its
    data sit in its records, not in lists): the owners add those records on top of the entries they pull."""
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(4)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    modes = [cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, cabi.MODE_FAST, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED]
    Gs = [Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=m) for m in modes]
    for G in Gs:
        G._ensure(2, table, p1, p2, l2d, l3d)
    blobs = [G.peer_export() for G in Gs]
    for r, G in enumerate(Gs):
        G.peer_attach(3, r, blobs)
    for fi, c in enumerate(files):
        Gs[fi % 2].accum_ods(c, iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
    assert cabi.load().gritas_b200_tiled(Gs[1].handle) != 1          # undecided / direct: nothing parked on rank 1
    for G in Gs:
        G.sync_no_flush()
    for G in Gs:
        G.export_planes()
    out = Gs[0].alloc_grids()
    for a in out.values():
        a[...] = -777.0
    for G in Gs:
        G.norm_sharded_into(out)
    for G in Gs:
        G.close()
    ref, _ = H.oracle_run(cfg, files, 2, [H.filter_in(c, 101) for c in files])
    H.compare(out, ref, l2d, l3d, 2, exact=False, tol=1e-6)


def test_sharded_norm_surface_grids_only():
    cfg = synth.get_config("cfg1", 0.03)
    cfg.im, cfg.jm = 144, 91
    cfg.kt_list = [1, 2, 3, -44]
    cfg.kx_lists = [[181, 187], [181], [180], [120]]
    files = []
    for i in range(3):
        c = synth.make_file(synth.get_config("cfg1", 0.03), i)
        rng = np.random.default_rng(200 + i)
        row = rng.integers(0, 4, len(c["lat"]))
        c["kt"] = np.asarray([1, 2, 3, 44], np.int32)[row]
        c["kx"] = np.asarray([181, 181, 180, 120], np.int32)[row]
        files.append(c)
    out, slabs, tiled, (l2d, l3d) = _run(cfg, files, 1, 3, cabi.MODE_FAST | cabi.FLAG_FORCE_TILED, 1)
    assert (l2d, l3d) == (4, 0)
    ref, _ = H.oracle_run(cfg, files, 1, [H.filter_in(c, 1) for c in files])
    H.compare(out, ref, l2d, l3d, 1, exact=False, tol=1e-6)
