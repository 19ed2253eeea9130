#!/usr/bin/env python3
"""Turns one `ncu --set full --import-source on` report into the short text summary kept under profiles/:
the launch's key metrics (raw page) and the source lines that collect the most warp-stall samples.
   usage: ncu_summary.py report.ncu-rep [algorithmic_bytes] > profiles/<name>_summary.txt"""
import collections
import csv
import json
import os
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration (us)"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block (KB)"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: shared memory (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps)"),
    ("dram__bytes_read.sum", "dram bytes read (MB)"), ("dram__bytes_write.sum", "dram bytes written (MB)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (% of peak)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (% of peak)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "barrier", "branch_resolving",
          "lg_throttle", "mio_throttle", "no_instruction", "dispatch_stall", "membar", "sleeping"]


def ncu(*args):
    return subprocess.run(["ncu", "-i"] + list(args), capture_output=True, text=True).stdout.splitlines()


def main():
    rep = sys.argv[1]
    alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv")))
    d = dict(zip(rows[0], rows[2]))
    unit = dict(zip(rows[0], rows[1]))
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3,
             "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):      # -> us / MB
        if k in d and unit.get(k) in scale:
            d[k] = "%.3f" % (float(d[k].replace(",", "")) * scale[unit[k]])
    print("report:", os.path.basename(rep))
    print("kernel:", d.get("Kernel Name", "?")[:150])
    for k, label in KEYS:
        if k in d:
            print("  %-46s %s" % (label, d[k]))
    try:
        dur_us = float(d["gpu__time_duration.sum"])
        traffic = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * 1e6
        print("  %-46s %.1f" % ("dram traffic / duration (GB/s, under ncu)", traffic / dur_us / 1e3))
        if alg:
            print("  %-46s %.1f  (algorithmic bytes %d; traffic / algorithmic = %.3f)" %
                  ("algorithmic bytes / duration (GB/s, under ncu)", alg / dur_us / 1e3, int(alg), traffic / alg))
        peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
        if os.path.exists(peaks) and alg:
            hbm = float(json.load(open(peaks))["hbm_gbs"])
            print("  %-46s %.3f  (peak %.1f GB/s, MEASURED_PEAKS.json; ncu's duration is cold-cache and serialised)" %
                  ("fraction of the measured HBM peak", alg / dur_us / 1e3 / hbm, hbm))
    except (KeyError, ValueError):
        pass
    print("  warp stalls per issued instruction:")
    for st in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
        if k in d:
            print("    %-22s %s" % (st, d[k]))
    # ---- source lines by warp-stall samples ---------------------------------------------------------------------
    out = ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda")
    cur, hdr, last = None, None, None
    agg, ex, src = collections.Counter(), collections.Counter(), {}
    for row in csv.reader(out):
        if not row:
            continue
        if row[0] == "File Path":
            cur = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row
            isamp, iex = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            continue
        if hdr is None:
            continue
        if row[0] == "":
            if last is not None:
                try:
                    agg[last] += int(row[isamp])
                    ex[last] += int(row[iex])
                except ValueError:
                    pass
            continue
        try:
            last = (cur, int(row[0]))
            src[last] = row[1]
        except ValueError:
            pass
    tot, totex = sum(agg.values()), sum(ex.values())
    if tot:
        print("  source lines by warp-stall samples (%d samples, %d warp instructions):" % (tot, totex))
        for k, v in agg.most_common(14):
            print("    %5.1f%% of samples, %5.1f%% of instructions  %s:%d  %s" %
                  (100.0 * v / tot, 100.0 * ex[k] / max(1, totex), k[0], k[1], src.get(k, "").strip()[:110]))


if __name__ == "__main__":
    main()
