"""Summarise an `ncu --page raw --csv` export: one line per captured launch with the metrics the roofline discussion uses.
usage: python tools/summarise_ncu_raw.py gpurun_out/rXX_streaming_raw.csv "header note" > profiles/rXX_streaming_ncu.csv"""
import csv
import re
import sys

path, note = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("smsp__inst_executed.sum", "inst"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_static", "smem_static"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("launch__occupancy_limit_registers", "occ_lim_regs"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall_long_sb_pct"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio_throttle"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle")]
idx = [(hdr.index(a), b) for a, b in want if a in hdr]
if note:
    print("# " + note)
print(",".join(b + (f" [{units[i]}]" if units[i] else "") for i, b in idx))
for r in rows[2:]:
    out = []
    for i, b in idx:
        v = r[i]
        if b == "kernel":
            v = re.sub(r"\(.*", "", v).strip().replace(",", ";")
        out.append(v.replace(",", ""))
    print(",".join(out))
