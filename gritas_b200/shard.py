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
