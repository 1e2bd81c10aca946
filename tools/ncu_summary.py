"""Key per-launch metrics of every kernel in an ncu report (used for profiles/*_summary.md).
usage: python tools/ncu_summary.py <report.ncu-rep> [...]"""
import csv, subprocess, sys
WANT = [("duration us", "gpu__time_duration.sum"), ("DRAM read MB", "dram__bytes_read.sum"), ("DRAM write MB", "dram__bytes_write.sum"),
        ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue slots %", "smsp__issue_active.avg.pct"), ("eligible warps/cycle", "smsp__warps_eligible.avg.per_cycle_active"),
        ("warp instructions", "smsp__inst_executed.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("regs/thread", "launch__registers_per_thread"), ("dyn smem KB", "launch__shared_mem_per_block_dynamic"),
        ("SM clock GHz", "sm__cycles_elapsed.avg.per_second")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"== {rep}: {name[:110]}")
        for label, key in WANT:
            for i, h in enumerate(hdr):
                if h == key or (h.startswith(key) and key.endswith(".pct")):
                    print(f"   {label:22s} {r[i]:>14s} {units[i]}")
                    break
