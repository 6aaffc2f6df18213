"""Summarise an .ncu-rep (read here, no GPU): one block per launch with the metrics that matter for HBM-bound kernels.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--csv out.csv]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cyc_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), short, m) for m, short in WANT if m in hdr]
    kn = hdr.index("Kernel Name")
    table = [["kernel"] + [f"{s} [{units[i]}]" for i, s, _ in cols]]
    for r in rows[2:]:
        table.append([r[kn][:70]] + [r[i] for i, _, _ in cols])
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as f:
            csv.writer(f).writerows(table)
    for r in table[1:]:
        print("==", r[0])
        for h, v in zip(table[0][1:], r[1:]):
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            print(f"   {h:42s} {v}")


if __name__ == "__main__":
    main()
