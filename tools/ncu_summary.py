#!/usr/bin/env python
"""Print the headline metrics of every kernel in an .ncu-rep (ncu --page raw --csv), and optionally the hottest
source lines.  usage: tools/ncu_summary.py report.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "smsp__inst_executed_op_global_ld.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name[:100])
        for k in WANT:
            if k in hdr:
                i = hdr.index(k)
                print("  %-85s %s %s" % (k, r[i], units[i]))
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                             text=True).stdout
        cur, hdr, agg = "?", None, {}
        for r in csv.reader(io.StringIO(out)):
            if not r:
                continue
            if r[0] == "File Path":
                cur = r[1].split("/")[-1]
            elif r[0] == "Line No":
                hdr = r
            elif hdr and len(r) == len(hdr) and r[0].isdigit():
                key = (cur, int(r[0]))
                samp = int(r[hdr.index("# Samples")] or 0)
                inst = int(r[hdr.index("Instructions Executed")] or 0)
                thr = int(r[hdr.index("Thread Instructions Executed")] or 0)
                a0 = agg.setdefault(key, [0, 0, 0, r[1]])
                a0[0] += samp
                a0[1] += inst
                a0[2] += thr
        tot = sum(v[0] for v in agg.values()) or 1
        toti = sum(v[1] for v in agg.values()) or 1
        print("-- source lines by instructions executed (stall samples %, warp instructions %, avg threads)")
        for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
            print("  %5.1f%% %5.1f%% %5.1f  %s:%-4d %s" % (100.0 * v[0] / tot, 100.0 * v[1] / toti, v[2] / max(v[1], 1), f, ln, v[3].strip()[:110]))


if __name__ == "__main__":
    main()
