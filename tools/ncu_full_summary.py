"""Summary of an `ncu --set full` report: for every profiled launch the handful of counters the roofline argument needs
(duration, DRAM bytes, pipe utilisation, issue rate, occupancy, top stall reasons).

    python tools/ncu_full_summary.py gpurun_out/<name>.ncu-rep > profiles/<name>_summary.txt

Reads the report through `ncu -i ... --page raw --csv` (the CLI in this image; no GPU needed)."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path: str) -> None:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(head)}
    for r in rows[2:]:
        print(f"---- {r[col['Kernel Name']][:150]}   grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEEP:
            if k in col:
                print(f"  {k:85s} {r[col[k]]:>16s} {units[col[k]]}")
        stalls = []
        for name, i in col.items():
            if name.startswith(STALL) and name.endswith("_per_warp_active.pct") is False and name.endswith(".ratio"):
                try:
                    stalls.append((float(r[i]), name[len(STALL):-len(".ratio")]))
                except ValueError:
                    pass
        if stalls:
            print("  warp stall reasons (warps stalled per issue-active cycle), top 5:")
            for v, name in sorted(stalls, reverse=True)[:5]:
                print(f"    {name:40s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
