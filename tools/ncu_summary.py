"""Condense an `ncu -i X.ncu-rep --page raw --csv` export into the short text files kept under profiles/.

usage: python tools/ncu_summary.py raw.csv [launch_index ...] > profiles/rN_<kernel>_summary.txt
Only reads the CSV; never runs anything.
"""
import csv
import sys

KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum lts__t_bytes.sum
sm__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active launch__registers_per_thread launch__grid_size launch__block_size
launch__shared_mem_per_block_dynamic smsp__inst_executed.sum lts__t_sector_hit_rate.pct
smsp__issue_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__warps_eligible.avg.per_cycle_active launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
launch__occupancy_limit_warps sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active
l1tex__data_bank_conflicts_pipe_lsu.sum smsp__sass_inst_executed_op_shared_ld.sum smsp__sass_inst_executed_op_shared_st.sum
smsp__sass_inst_executed_op_global_ld.sum smsp__sass_inst_executed_op_global_st.sum""".split()


def main():
    path = sys.argv[1]
    wanted = [int(a) for a in sys.argv[2:]]
    rows = list(csv.reader(open(path, newline="")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    head, units, body = rows[start], rows[start + 1], rows[start + 2:]
    col = {name: i for i, name in enumerate(head)}
    print(f"# ncu --set full --clock-control none, exported with --page raw --csv ({path.split('/')[-1]})")
    for n, r in enumerate(body):
        if wanted and n not in wanted:
            continue
        print(f"## launch {n}")
        print(f"{'Kernel Name':96s}{r[col['Kernel Name']]}")
        for k in KEEP:
            if k in col:
                print(f"{k:96s}{r[col[k]]} {units[col[k]]}")
        print()


if __name__ == "__main__":
    main()
