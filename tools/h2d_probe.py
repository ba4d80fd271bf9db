#!/usr/bin/env python
"""Host <-> device copy ceiling of this box with N concurrent ranks (what bounds the end-to-end numbers).

Launch like the bench:  python -m torch.distributed.run --nproc-per-node N tools/h2d_probe.py  (or plain python for N=1).
Every rank copies pinned 1 GiB buffers to and from its GPU: alone (rank 0 only), all ranks together H2D, all together
D2H, and H2D + D2H at once on two streams.  Rank 0 prints one JSON line; bench.py's e2e can be read as a fraction of it.
Also reports where the rank's pinned memory and CPU affinity sit (NUMA), since that is what differs between ranks.
"""
import json
import os
import time

import torch
import torch.distributed as dist


def numa_of_cpus():
    try:
        cpus = sorted(os.sched_getaffinity(0))
        nodes = set()
        for n in os.listdir("/sys/devices/system/node"):
            if n.startswith("node"):
                lst = open(f"/sys/devices/system/node/{n}/cpulist").read().strip()
                for part in lst.split(","):
                    a, _, b = part.partition("-")
                    lo, hi = int(a), int(b or a)
                    if any(lo <= c <= hi for c in cpus):
                        nodes.add(int(n[4:]))
        return {"cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})", "numa_nodes": sorted(nodes)}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 1 << 30
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=4, only_rank0=False):
        fn()
        sync_all()
        t0 = time.perf_counter()
        if not only_rank0 or rank == 0:
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        sync_all()
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1 and not only_rank0:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def h2d():
        dev.copy_(host, non_blocking=True)

    def d2h():
        host2.copy_(dev2, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            dev.copy_(host, non_blocking=True)
        with torch.cuda.stream(s2):
            host2.copy_(dev2, non_blocking=True)

    gb = nbytes / 1e9
    res = {"n_ranks": world, "bytes_per_copy": nbytes}
    res["h2d_alone_gbs"] = gb / timed(h2d, only_rank0=True)
    res["d2h_alone_gbs"] = gb / timed(d2h, only_rank0=True)
    t = timed(h2d)
    res["h2d_all_gbs_per_rank"], res["h2d_all_gbs_total"] = gb / t, gb * world / t
    t = timed(d2h)
    res["d2h_all_gbs_per_rank"], res["d2h_all_gbs_total"] = gb / t, gb * world / t
    t = timed(both)
    res["duplex_all_gbs_per_rank_each_way"], res["duplex_all_gbs_total_each_way"] = gb / t, gb * world / t
    where = [None] * world
    me = {"rank": rank, "gpu": local, **numa_of_cpus()}
    if world > 1:
        dist.all_gather_object(where, me)
    else:
        where = [me]
    res["ranks"] = where
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
