#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report as markdown: python tools/ncu_summary.py rep.ncu-rep [title]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (all)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
]


def main(path, title=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    print("# %s\n" % (title or path))
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("## %s  grid %s block %s\n" % (rec.get("Kernel Name", "?")[:90], rec.get("Grid Size"), rec.get("Block Size")))
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in rec:
                print("| %s (`%s`) | %s | %s |" % (label, k, rec[k], u.get(k, "")))
        print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
