"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
usage: python tools/ncu_summary.py <tag> [launches.csv] [prof.ncu-rep]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def launches(path, out):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in csv.DictReader(io.StringIO("".join(rows))):
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(d["Metric Unit"], 1.0)
        agg[d["Kernel Name"][:70]][0] += 1
        agg[d["Kernel Name"][:70]][1] += v
    tot = sum(v[1] for v in agg.values())
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    out.write("%-72s %6s %12s %10s %7s\n" % ("kernel", "n", "total_ms", "avg_us", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("%-72s %6d %12.3f %10.1f %6.1f%%\n" % (n, c, t / 1e6, t / c / 1e3, 100 * t / tot))


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    hdr, units = r[0], r[1]
    out.write("# ncu --set full --clock-control none: selected metrics per captured launch\n")
    for row in r[2:]:
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                out.write("%-90s %s %s\n" % (k, row[i], units[i]))
        out.write("\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] != "-":
        with open("profiles/%s_launches.txt" % tag, "w") as f:
            launches(sys.argv[2], f)
    if len(sys.argv) > 3:
        with open("profiles/%s_ncu_full.txt" % tag, "w") as f:
            full(sys.argv[3], f)
