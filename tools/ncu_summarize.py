"""Turns `ncu -i X.ncu-rep --page raw --csv` into the small JSON summaries committed under profiles/."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem",
    "sm__cycles_elapsed.max": "sm_cycles",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed": "dram_cycles_active_pct",
}


def main(rep, out, commit, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        e = {"kernel": d.get("Kernel Name", "")[:160]}
        for k, name in KEEP.items():
            if k in d and d[k] != "":
                try:
                    e[name] = float(d[k].replace(",", ""))
                except ValueError:
                    e[name] = d[k]
                u = units[hdr.index(k)]
                if u:
                    e[name + "_unit"] = u
        res.append(e)
    # normalise the byte / time units ncu picks per column
    for e in res:
        for f in ("dram_read_bytes", "dram_write_bytes"):
            u = e.get(f + "_unit", "byte").lower()
            mul = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)
            if f in e:
                e[f] = e[f] * mul
                e[f + "_unit"] = "byte"
        u = e.get("duration_ns_unit", "ns").lower()
        mul = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(u, 1)
        if "duration_ns" in e:
            e["duration_ns"] *= mul
            e["duration_ns_unit"] = "ns"
        if "dram_read_bytes" in e and "dram_write_bytes" in e:
            e["traffic_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
            if e.get("duration_ns"):
                e["dram_gbs"] = e["traffic_bytes"] / e["duration_ns"]
    top = max(res, key=lambda e: e.get("duration_ns", 0)) if res else {}
    doc = {"source": rep.split("/")[-1], "commit": commit, "note": note, "launches": res,
           "traffic_bytes_per_launch": top.get("traffic_bytes"), "duration_ns": top.get("duration_ns")}
    json.dump(doc, open(out, "w"), indent=1)
    print(out, json.dumps({k: top.get(k) for k in ("kernel", "duration_ns", "traffic_bytes", "dram_gbs", "warps_active_pct", "l1_hit_pct", "sm_throughput_pct")}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "", sys.argv[4] if len(sys.argv) > 4 else "")
