"""Prints the stall-reason / throughput lines of an `ncu --page raw --csv` export (development tool)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print(d.get("Kernel Name", "")[:100])
    st = [(k, float(d[k].replace(",", ""))) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and d.get(k)]
    for k, v in sorted(st, key=lambda kv: -kv[1])[:6]:
        print(f"  {v:9.3f} {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
              "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        if k in d:
            print(f"  {k} = {d[k]}")
