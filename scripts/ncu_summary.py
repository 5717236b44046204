"""Summarise .ncu-rep files (ncu -i ... --page raw --csv) into the handful of metrics DESIGN.md quotes."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct2"),
    ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_elapsed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active"),
    ("smsp__inst_executed.sum", "inst"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        print(f"==== {path}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print("--", d.get("Kernel Name", "?")[:110])
            for k, short in WANT:
                if k in d and d[k] not in ("", "no data"):
                    print(f"   {short:18s} {d[k]} {u[k]}")


if __name__ == "__main__":
    main()
