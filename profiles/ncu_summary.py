#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`): headline counters + instruction/sample hot spots
by source line.  Usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'inst_executed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed_op_tma_ld.sum']


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:110])
        for w in WANT:
            if w in hdr:
                print(f"  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}")
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v >= 0.05:
                    print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} = {v:.2f} per issue")
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    agg, samp, text = collections.Counter(), collections.Counter(), {}
    hdr, fname = None, ""
    for r in rows:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]; hdr = None; continue
        if r and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < 10:
            continue
        try:
            ln = int(r[0]); n = int(r[hdr.index("Instructions Executed")] or 0); s = int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            continue
        agg[(fname, ln)] += n; samp[(fname, ln)] += s; text[(fname, ln)] = r[1][:96]
    tot, ts = sum(agg.values()) or 1, sum(samp.values()) or 1
    print(f"== source hot spots (warp instructions {tot}, stall samples {ts})")
    keys = sorted(agg, key=lambda k: -(agg[k] / tot + samp[k] / ts))[:top]
    for k in keys:
        print(f"  {agg[k] / tot * 100:5.1f}% inst {samp[k] / ts * 100:5.1f}% samples  {k[0]}:{k[1]:<4d} {text[k]}")


if __name__ == "__main__":
    main()
