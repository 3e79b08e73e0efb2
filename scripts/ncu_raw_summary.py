"""Key metrics of an ncu report: python scripts/ncu_raw_summary.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
for v in rows[2:]:
    print("==", v[h.index("Kernel Name")][:90])
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
    for i, name in enumerate(h):
        if name in want:
            print(f"  {name} = {v[i]}")
    st = [(float(v[i]), name.split("issue_stalled_")[1].split("_per_")[0]) for i, name in enumerate(h)
          if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio") and v[i]]
    print("  stalls/issue:", ", ".join(f"{n}={x:.2f}" for x, n in sorted(st, reverse=True)[:8]))
