"""Print a few named metrics per launch from an `ncu --page raw --csv` export.  Usage: python profiles/ncu_pick.py raw.csv [substr ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pick = sys.argv[2:] or ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
                        "sm__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size"]
name = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
for r in data:
    print((r[name][:60] if name is not None else "?"))
    for i, h in enumerate(hdr):
        if any(h == p or (p.endswith("*") and h.startswith(p[:-1])) for p in pick):
            print(f"    {h} [{units[i]}] = {r[i]}")
