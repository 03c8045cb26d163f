#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv   > profiles/<name>.md   (per-kernel time shares)
  python profiles/summarize.py full     gpurun_out/prof.ncu-rep   > profiles/<name>.md   (--set full key metrics)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0][:70]


def launches(path, anchor=None, lo=0, hi=None):
    """anchor: only the launches from the lo-th to the hi-th occurrence of a kernel whose name contains `anchor` (the
    timed steps of bench.py: every step starts with the same kernel)"""
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    if anchor:
        ki = hdr.index("Kernel Name")
        # one row per launch here (a single metric), so occurrences can be counted on the rows directly
        occ = [i for i, r in enumerate(rows[1:], 1) if anchor in r[ki]]
        a = occ[lo]
        b = occ[hi] if hi is not None and hi < len(occ) else len(rows)
        rows = [hdr] + rows[a:b]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ui = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)   # -> us
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches (cold-cache, serialised under ncu: compare shares)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"### `{short(r[hdr.index('Kernel Name')])}`  (ID {r[0]})\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                print(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        rd = float(r[hdr.index('dram__bytes_read.sum')].replace(",", ""))
        wr = float(r[hdr.index('dram__bytes_write.sum')].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        tr = rd * scale[units[hdr.index('dram__bytes_read.sum')]] + wr * scale[units[hdr.index('dram__bytes_write.sum')]]
        print(f"| traffic = dram read + write | {tr / 1e9:.3f} | GB |\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches" and len(sys.argv) > 3:
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
