"""Summarise one-kernel `ncu --set full` reports: python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt"""
import csv, subprocess, sys
KEYS = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active dram__throughput.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed l1tex__throughput.avg.pct_of_peak_sustained_elapsed smsp__inst_executed.sum
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio smsp__cycles_active.avg sm__cycles_elapsed.max""".split()
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("Kernel Name =", r[hdr.index("Kernel Name")])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} [{units[i]}] = {r[i]}")
