#!/usr/bin/env python
"""Summarise ncu output for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum)
into per-kernel totals and shares; (2) a --set full .ncu-rep into the handful of metrics the
roofline discussion uses.  Usage:
    python scripts/ncu_summary.py launches <launches.csv>
    python scripts/ncu_summary.py full <report.ncu-rep> [kernel-regex]
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot = {}
    for r in rows[start + 1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        name = re.sub(r"\(.*", "", r[ik]).replace("void <unnamed>::", "")
        t = tot.setdefault(name, [0, 0.0]); t[0] += 1; t[1] += v
    total = sum(v[1] for v in tot.values())
    print("%-60s %8s %12s %8s" % ("kernel", "launches", "total_us", "share"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.1f %7.1f%%" % (k[:60], v[0], v[1], 100 * v[1] / total))
    print("%-60s %8d %12.1f" % ("TOTAL", sum(v[0] for v in tot.values()), total))


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pattern and not re.search(pattern, name):
            continue
        print("kernel:", name)
        for k in KEYS:
            if k in hdr:
                print("  %-70s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = [(h, float(r[hdr.index(h)])) for h in hdr if h.startswith("smsp__average_warp") is False and
                  "issue_stalled" in h and h.endswith("_per_warp_active.pct") and r[hdr.index(h)] not in ("", "n/a")]
        for h, v in sorted(stalls, key=lambda kv: -kv[1])[:8]:
            print("  stall %-64s %.1f %%" % (h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", ""), v))
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
