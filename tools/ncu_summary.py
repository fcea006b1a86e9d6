"""Prints the judged metrics of every kernel in an .ncu-rep (read on the CPU box with `ncu -i`)."""
import csv, subprocess, sys
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
for r in rows[2:]:
    print("---")
    for w in WANT:
        if w in H:
            i = H.index(w)
            print(f"{w} = {r[i]} {rows[1][i]}")
