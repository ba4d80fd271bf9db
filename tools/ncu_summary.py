#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (raw metrics page + top stall lines of the source page).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt   (needs `ncu` on PATH, no GPU)"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    ik = hdr.index("Kernel Name")
    print(f"# {rep}: {len(raw) - 2} profiled launches (ncu --set full --clock-control none; cold caches, serialised)")
    seen = set()
    for r in raw[2:]:
        name = r[ik]
        if name in seen:
            continue
        seen.add(name)
        print(f"\n== {name[:100]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {r[i]:>18s} {units[i]}")
    src = page(rep, "source")
    kern, hdr, data, names = -1, None, {}, {}
    for r in src:
        if r and r[0] == "Kernel Name":
            kern += 1
            data[kern] = []
            names[kern] = r[1]
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) > 10 and kern >= 0:
            data[kern].append(r)
    if hdr is None:
        return
    i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    done = set()
    for k, out in data.items():
        if names[k] in done or not out:
            continue
        done.add(names[k])
        tot = sum(int(r[i_s]) for r in out)
        agg = {s: sum(int(r[hdr.index(s)]) for r in out) for s in stalls}
        print(f"\n== stalls: {names[k][:90]}\n  samples {tot}: " +
              ", ".join(f"{s[6:]} {100 * v / max(tot, 1):.0f}%" for s, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
        cls = collections.Counter()
        for r in out:
            m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_]+)", r[i_src])
            cls[m.group(2) if m else "?"] += int(r[i_ex])
        t = sum(cls.values())
        print("  warp instructions %d: " % t + ", ".join(f"{op} {100 * c / t:.0f}%" for op, c in cls.most_common(8)))
        for r in sorted(out, key=lambda r: -int(r[i_s]))[:8]:
            st = {s[6:]: int(r[hdr.index(s)]) for s in stalls if int(r[hdr.index(s)]) > 0}
            top = sorted(st.items(), key=lambda x: -x[1])[:2]
            print(f"    {r[i_s]:>5s} samples  {r[i_src].strip()[:64]:64s} {top}")


if __name__ == "__main__":
    main(sys.argv[1])
