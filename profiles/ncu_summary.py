#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few counters the roofline needs."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"]
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units = r[0], r[1]
    def col(name):
        for i, h in enumerate(hdr):
            if h.endswith(name):
                return i
        return None
    print("| kernel | " + " | ".join(w.split(".")[0] + "." + w.split(".")[-1] if "." in w else w for w in WANT) + " |")
    for row in r[2:]:
        name = row[col("Kernel Name")][:60]
        vals = []
        for w in WANT:
            i = col(w)
            vals.append("-" if i is None else f"{row[i]} {units[i]}")
        print(f"| {name} | " + " | ".join(vals) + " |")
if __name__ == "__main__":
    main(sys.argv[1])
