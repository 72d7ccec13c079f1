"""Summarise an .ncu-rep (read here, no GPU needed) into a small CSV/markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep profiles/r1_ncu_full_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "L2_hit_%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_%"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    # what bounds a gather / rows / tile kernel: L2->SM sectors, shared-memory bank conflicts, stall reasons
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->SM_sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "L1_ld_sectors"),
    ("l1tex__t_sector_hit_rate.pct", "L1_hit_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("lts__t_sectors_op_red.sum", "L2_red_sectors"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), n) for m, n in METRICS if m in hdr]
    kname = hdr.index("Kernel Name")
    lines = ["| # | kernel | " + " | ".join(n for _, n in cols) + " |", "|---|---|" + "---|" * len(cols)]
    for i, r in enumerate(data):
        name = r[kname].split("(")[0].replace("void ", "")
        vals = []
        for c, n in cols:
            v = r[c]
            try:
                f = float(v)
                v = f"{f:.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {units[c]}".strip())
        lines.append(f"| {i} | `{name}` | " + " | ".join(vals) + " |")
    with open(out, "w") as fh:
        fh.write(f"Source: `{rep}` (ncu --set full --clock-control none; per-launch values, cold caches, kernel replay)\n\n")
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
