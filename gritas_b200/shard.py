"""Partitioning of synoptic-time files over ranks (one process per GPU).

The reference is single-process; its only concurrency is independent background gritas.x runs
(src/Components/gritas/GIO/run_sat_gritas.csh:106-127).  Files are independent units and the
accumulators are commutative sums, so file f goes to rank f mod world and the
partial count / sum / sum-of-squares grids are summed once before GrITAS_Norm.
"""
from __future__ import annotations


def files_for_rank(nfiles: int, rank: int, world: int) -> list:
    """Indices of the files rank `rank` of `world` accumulates (round-robin, balanced to +-1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, nfiles, world))


def owner_of(file_index: int, world: int) -> int:
    return file_index % world


def slab_rows(im: int, jm: int, km: int, l2d: int, l3d: int, rank: int, world: int):
    """Rows (jl = j + jm*l, 0-based, half-open) of the 3-D and of the 2-D grids whose statistics `rank` finalizes and writes in a
    sharded Norm when nothing is known about the data yet: the records [nrec*r/N, nrec*(r+1)/N) snapped down to whole rows --
    the same partition gritas_b200_slab returns before the first sharded Norm (which then balances the slabs by parked
    entries).  Rows are never split, every row belongs to exactly one rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    nbox3, nbox2 = im * jm * km * l3d, im * jm * l2d
    nrec = nbox3 + nbox2
    row3 = im * km

    def bound(r):
        if r <= 0:
            return 0
        if r >= world:
            return nrec
        x = nrec * r // world
        if x <= nbox3:
            return x // row3 * row3
        return nbox3 + (x - nbox3) // im * im
    lo, hi = bound(rank), bound(rank + 1)
    return (min(lo, nbox3) // row3, min(hi, nbox3) // row3, (max(lo, nbox3) - nbox3) // im, (max(hi, nbox3) - nbox3) // im)
