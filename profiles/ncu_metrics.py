#!/usr/bin/env python
"""Print the metrics that matter from an .ncu-rep (via `ncu -i … --page raw --csv`)."""
import csv
import io
import subprocess
import sys

KEYS = """gpu__time_duration.sum launch__grid_size launch__registers_per_thread
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sectors.sum lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sectors_op_red.sum lts__t_sectors_op_atom.sum lts__t_requests_srcunit_tex_op_red.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__t_sector_hit_rate.pct l1tex__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_fp64.sum smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_drain_per_issue_active.ratio
smsp__average_warps_issue_stalled_membar_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio""".split()


def main(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    for d in data:
        if pattern and pattern not in d[kn]:
            continue
        print("==", d[kn][:60])
        for k in KEYS:
            if k in hdr:
                j = hdr.index(k)
                print(f"  {k:85s} {d[j]:>16s} {units[j]}")
    if len(sys.argv) > 3:
        for h in hdr:
            if sys.argv[3] in h:
                print(h, [d[hdr.index(h)] for d in data])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None)
