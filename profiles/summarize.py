"""Summaries of ncu output kept under profiles/: `python profiles/summarize.py launches X.csv` / `raw X.ncu-rep`."""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in d.values())
    print("kernel,launches,mean_us,share_pct")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("%s,%d,%.2f,%.1f" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[ki])
        for m in METRICS:
            if m in hdr:
                print("   %-75s %s %s" % (m, r[hdr.index(m)], units[hdr.index(m)]))
        for i, h in enumerate(hdr):
            if "stall" in h and h.endswith("_per_warp_active.pct"):
                try:
                    if float(r[i]) > 5:
                        print("   %-75s %s" % (h, r[i]))
                except ValueError:
                    pass


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
