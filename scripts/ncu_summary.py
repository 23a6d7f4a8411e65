#!/usr/bin/env python
"""Condense ncu output into the small tracked files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof_syrk_r1.ncu-rep profiles/r1_syrk_full.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size",
    "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        us = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else (v * 1e6 if unit == "s" else v))
        agg.setdefault(name, []).append(us)
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`)\n\n")
        f.write("source: `%s` (%d launches, %.1f ms of kernel time; cold-cache, serialised: compare SHARES)\n\n" % (
            src, sum(len(v) for v in agg.values()), tot / 1000))
        f.write("| kernel | launches | total us | share | mean us | min us | max us |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.1f | %.1f%% | %.2f | %.2f | %.2f |\n" % (
                k, len(v), sum(v), 100 * sum(v) / tot, sum(v) / len(v), min(v), max(v)))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of `%s`\n\n" % src)
        ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---|" * len(data) + "\n")
        f.write("| kernel | | " + " | ".join("`%s`" % r[ki] for r in data) + " |\n")
        f.write("| grid / block | | " + " | ".join("%s / %s" % (r[gi], r[bi]) for r in data) + " |\n")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                f.write("| %s | %s | %s |\n" % (k, units[i], " | ".join(r[i] for r in data)))
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
