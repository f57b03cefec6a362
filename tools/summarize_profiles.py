#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the text summaries committed under profiles/.

  python tools/summarize_profiles.py launches gpurun_out/launches_r1d.csv "<header line>"  > profiles/r1d_launches_summary.txt
  python tools/summarize_profiles.py ncu gpurun_out/prof_graph_r1d.ncu-rep "<header line>" > profiles/r1d_graph_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum dram__throughput.avg.pct_of_peak_sustained_elapsed
launch__registers_per_thread launch__occupancy_limit_registers launch__occupancy_limit_shared_mem launch__waves_per_multiprocessor
sm__warps_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active smsp__issue_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
lts__t_sector_hit_rate.pct l1tex__t_sector_hit_rate.pct l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio""".split()


def launches(path, header):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = r["Kernel Name"]
        tot[k][0] += 1
        tot[k][1] += ms
    total = sum(v[1] for v in tot.values())
    print(header)
    print("(cold-cache, serialised launches: compare SHARES, not absolute times)\n")
    print("%-92s %8s %10s %7s" % ("kernel", "launches", "total ms", "share"))
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %8d %10.3f %6.1f%%" % (k[:92], n, ms, 100 * ms / total))


def ncu(path, header):
    out = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(header)
    for vals in rows[2:]:
        print("-----")
        d = dict(zip(hdr, zip(units, vals)))
        for name in ("Kernel Name", "Block Size", "Grid Size"):
            print("%s [] = %s" % (name, d[name][1]))
        for m in METRICS:
            if m in d:
                print("%s [%s] = %s" % (m, d[m][0], d[m][1]))


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
