"""Print the roofline-relevant raw metrics of every kernel in an .ncu-rep (run here, no GPU needed).
Usage: python scripts/ncu_brief.py gpurun_out/x.ncu-rep [more.ncu-rep ...]"""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TU = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1}


def main() -> None:
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            print(f"\n## {rep}  --  {d.get('Kernel Name', ('?', ''))[0][:150]}")
            for w in WANT:
                if w in d:
                    print(f"{w:88s} {d[w][0]:>18s} {d[w][1]}")
            try:
                rd = float(d["dram__bytes_read.sum"][0]) * UNIT.get(d["dram__bytes_read.sum"][1], 1)
                wr = float(d["dram__bytes_write.sum"][0]) * UNIT.get(d["dram__bytes_write.sum"][1], 1)
                dur = float(d["gpu__time_duration.sum"][0]) * TU.get(d["gpu__time_duration.sum"][1], 1)
                print(f"-> DRAM read {rd:.4e} B + write {wr:.4e} B = {rd + wr:.4e} B; {dur * 1e3:.3f} ms under ncu; "
                      f"{(rd + wr) / dur / 1e9:.0f} GB/s; warp-instructions {float(d['smsp__inst_executed.sum'][0]):.4e}")
            except (KeyError, ValueError):
                pass


if __name__ == "__main__":
    main()
