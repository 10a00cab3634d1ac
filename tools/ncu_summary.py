#!/usr/bin/env python
"""Summarise an ncu launch list (csv from --metrics gpu__time_duration.sum) per kernel, and optionally
the key counters of a .ncu-rep (--set full) for one kernel.
usage: tools/ncu_summary.py launches.csv [prof.ncu-rep]"""
import collections, csv, io, re, subprocess, sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = csv.DictReader(io.StringIO("".join(rows)))
tot = collections.OrderedDict()
for r in rd:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    t, n = tot.get(name, (0.0, 0))
    tot[name] = (t + float(r["Metric Value"].replace(",", "")), n + 1)
total = sum(t for t, _ in tot.values())
print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for k, (t, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{k[:70]:70s} {n:8d} {t/1e3:12.1f} {t/1e3/n:10.1f} {100*t/total:6.1f}%")
if len(sys.argv) > 2:
    out = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rd[0], rd[1], rd[2:]
    want = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__inst_executed_pipe_(alu|fma|fmaheavy|fp64|lsu|uniform|xu).sum$|"
                      r"sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active|sm__pipe_fma_cycles_active.avg.pct|smsp__inst_executed.sum$|sm__inst_executed.sum$|"
                      r"smsp__issue_active.avg.pct|sm__warps_active.avg.pct_of_peak|launch__registers_per_thread|launch__occupancy_limit|sm__throughput.avg.pct|"
                      r"smsp__cycles_active.avg$|sm__cycles_elapsed.max|smsp__thread_inst_executed_per_inst_executed.ratio|smsp__inst_executed_pipe_alu.sum$|"
                      r"l1tex__t_sector_hit_rate|lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_alu.avg.pct|sm__inst_issued.avg.pct|smsp__warp_issue_stalled.*_per_warp_active.pct")
    for v in vals:
        print("== kernel:", v[hdr.index("Kernel Name")][:90])
        for h, u, x in zip(hdr, units, v):
            if want.search(h):
                print(f"  {h:95s} {x:>16s} {u}")
