"""Extract the judged numbers from an `.ncu-rep` (run here, no GPU needed): `python benchmarks/ncu_extract.py file.ncu-rep`."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_op",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block", "launch__occupancy_limit",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__pcsamp_warps_issue_stalled",
        "smsp__average_warp", "smsp__warp_issue_stalled", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    print("kernel:", vals[head.index("Kernel Name")][:110])
    for h, u, v in zip(head, units, vals):
        if any(h.startswith(k) for k in KEYS):
            print(f"  {h:90s} {v:>16s} {u}")


if __name__ == "__main__":
    main()
