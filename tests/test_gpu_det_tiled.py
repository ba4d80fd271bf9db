"""Deterministic mode on the tiled path (k_list_obs<DET> + k_tile_det): every box is summed in input order -- call order,
then index within the call -- continuing from the sums it already has, so the statistics are BIT-IDENTICAL to the
sequential reference loop (m_gritas_grids.f:405-474) whatever order the tickets were handed out in."""
import numpy as np
import pytest

from gritas_b200 import cabi, shard, synth
from gritas_b200.m_gritas_grids import Grids
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu
DET = cabi.MODE_DETERMINISTIC


def _tiny_grid_cfg(scale, im=8, jm=5):
    """few tiles, long lists: every tile list is many chunks long (the multi-group path of k_tile_det)"""
    cfg = synth.get_config("cfg2", scale)
    cfg.im, cfg.jm = im, jm
    return cfg


@pytest.mark.parametrize("nr", [1, 2])
def test_long_lists_many_groups_bit_exact(nr):
    cfg = _tiny_grid_cfg(0.12)
    files = [synth.make_file(cfg, i) for i in range(3)]          # ~145 K obs each into 16.6 K boxes: ~20 K entries per tile
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    assert cfg.im * cfg.jm * cfg.km * l3d >= 4096
    fin = [H.filter_in(c, 1) for c in files]
    ref, rflags = H.oracle_run(cfg, files, nr, fin)
    out, gflags = H.gpu_run(cfg, files, nr, DET | cabi.FLAG_ASYNC, fin)
    assert all(np.array_equal(a, b) for a, b in zip(gflags, rflags))
    H.compare(out, ref, l2d, l3d, nr, exact=True, tol=0.0)
    out2, _ = H.gpu_run(cfg, files, nr, DET | cabi.FLAG_ASYNC, fin)
    for k in out:
        assert np.array_equal(out[k], out2[k], equal_nan=True), k  # run to run


def test_three_residual_series_bit_exact():
    cfg = _tiny_grid_cfg(0.05, 16, 9)
    files = [synth.make_file(cfg, i) for i in range(2)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    ref, _ = H.oracle_run(cfg, files, 3, norm=False)
    out, _ = H.gpu_run(cfg, files, 3, DET, norm=False)
    H.compare(out, ref, l2d, l3d, 3, exact=True, tol=0.0, raw=True)


def test_flush_between_calls_keeps_the_order():
    """Explicit flushes (lists -> records, in order) between the calls, then more calls, then Norm on lists + records."""
    cfg, nr = _tiny_grid_cfg(0.05, 24, 13), 2
    files = [synth.make_file(cfg, i) for i in range(5)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fin = [H.filter_in(c, 101) for c in files]
    ref, _ = H.oracle_run(cfg, files, nr, fin)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=DET)
    for i, c in enumerate(files):
        G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, nr), len(c["lat"]), nr, table, p1, p2, fin[i].copy(), l2d, l3d=l3d)
        if i in (1, 2):
            G.flush()
    out = G.alloc_grids()
    G.norm_into(out)
    again = G.alloc_grids()
    G.norm_into(again)                                           # Norm leaves the state as it is
    G.close()
    H.compare(out, ref, l2d, l3d, nr, exact=True, tol=0.0)
    for k in out:
        assert np.array_equal(out[k], again[k], equal_nan=True), k


def test_clustered_data_small_lists_flush_in_time(monkeypatch):
    """Radiosonde-site hot spots with short tile lists: the library flushes (in order) before a list fills up."""
    monkeypatch.setenv("GRITAS_B200_STAGE_SLOTS", str(1 << 21))
    cfg = synth.get_config("cfg5", 0.05)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(12)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fin = [H.filter_in(c, 101) for c in files]
    ref, _ = H.oracle_run(cfg, files, 2, fin)
    out, _ = H.gpu_run(cfg, files, 2, DET, fin)
    H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)


def test_one_box_beyond_a_chunk_is_reported():
    """More observations of ONE box parked than the tile kernel orders at once: complete sums, rc GRITAS_B200_ERR_ORDER."""
    cfg = _tiny_grid_cfg(0.01, 24, 13)
    c = synth.make_file(cfg, 0)
    n = 6000
    c = {k: (v[:n].copy() if isinstance(v, np.ndarray) else v) for k, v in c.items()}
    c["lat"][:] = 10.0
    c["lon"][:] = 20.0
    c["lev"][:] = 500.0
    c["kt"][:] = 44
    c["kx"][:] = 120
    c["qcexcl"][:] = 0
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    G = Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=DET)
    G.GrITAS_Accum(c["lat"], c["lon"], c["lev"], c["kx"], c["kt"], H.del_of(c, 1), n, 1, table, p1, p2, None, l2d, l3d=l3d)
    out = G.alloc_grids()
    with pytest.raises(RuntimeError, match="out of order"):
        G.norm_into(out)                                         # reported, and the arrays are filled all the same
    G.close()
    assert out["nobs_3d"].sum() == n
    ref, _ = H.oracle_run(cfg, [c], 1)
    H.compare(out, ref, l2d, l3d, 1, exact=False, tol=1e-9)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_norm_is_bit_identical_to_one_process(world):
    """Files f mod N on N handles, every call tagged with its global file index: the sharded Norm merges the ranks' entries
    of a box by sequence number -- the result is the single-process sequential sum, bit for bit, for any N."""
    cfg = synth.get_config("cfg2", 0.02)
    cfg.im, cfg.jm = 72, 46
    files = [synth.make_file(cfg, i) for i in range(7)]
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fin = [H.filter_in(c, 101) for c in files]
    ref, _ = H.oracle_run(cfg, files, 2, fin)
    Gs = [Grids(cfg.im, cfg.jm, cfg.km, cfg.plevs, mode=DET) for _ in range(world)]
    for G in Gs:
        G._ensure(2, table, p1, p2, l2d, l3d)
    blobs = [G.peer_export() for G in Gs]
    for r, G in enumerate(Gs):
        G.peer_attach(world, r, blobs)
    for r, G in enumerate(Gs):
        for fi in shard.files_for_rank(len(files), r, world):
            G.set_call_order(fi)
            G.accum_ods(files[fi], iopt=101, x_passive=synth.X_PASSIVE, x_pre_bad=synth.X_PRE_BAD)
    for G in Gs:
        G.sync_no_flush()
    out = Gs[0].alloc_grids()
    for G in Gs:
        G.norm_sharded_into(out)
    for G in Gs:
        G.close()
    H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)


@pytest.mark.parametrize("per_box", [33, 192, 193, 256, 257, 700, 1900])
def test_long_runs_in_few_boxes_bit_exact(per_box):
    """Hundreds of observations of the same box between two flushes (a radiosonde site): the run is ordered by a warp-level
    bitonic network instead of the quadratic rank count; lengths around the powers of two and just below a chunk."""
    cfg = _tiny_grid_cfg(0.004, 24, 13)
    files = []
    rng = np.random.default_rng(per_box)
    for i in range(2):
        c = synth.make_file(cfg, i)
        n = len(c["lat"])
        hot = rng.permutation(n)[: per_box // 2 + (per_box % 2) * (i == 0)]     # the hot box's entries are spread over the file
        c["lat"][hot], c["lon"][hot], c["lev"][hot] = 33.0, -77.0, 500.0
        c["kt"][hot], c["kx"][hot], c["qcexcl"][hot] = 44, 120, 0
        hot2 = np.setdiff1d(rng.permutation(n)[:300], hot)[:250]
        c["lat"][hot2], c["lon"][hot2], c["lev"][hot2] = 33.0, -77.0, 850.0        # a second long run in the same tile
        c["kt"][hot2], c["kx"][hot2], c["qcexcl"][hot2] = 44, 120, 0
        files.append(c)
    table, l2d, l3d, p1, p2 = H.setup(cfg)
    fin = [H.filter_in(c, 101) for c in files]
    ref, _ = H.oracle_run(cfg, files, 2, fin)
    assert ref.nobs_3d.max() >= per_box, ref.nobs_3d.max()
    out, _ = H.gpu_run(cfg, files, 2, DET | cabi.FLAG_ASYNC, fin)
    H.compare(out, ref, l2d, l3d, 2, exact=True, tol=0.0)
