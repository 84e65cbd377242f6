"""Per-source-line hot spots from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_source.py report.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
lines = txt.splitlines()
# several launches/files are concatenated; parse blocks that start with a header row containing "Source"
rows = []
hdr = None
cur_file = ""
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        continue
    if "Source" in row and "# Samples" in row:
        hdr = row
        continue
    if hdr is None:
        continue
    if len(row) != len(hdr):
        if len(row) <= 3:
            cur_file = ",".join(row)
        continue
    d = dict(zip(hdr, row))
    d["_file"] = cur_file
    rows.append(d)
def f(d, k):
    try:
        return float(d.get(k, "0").replace(",", "") or 0)
    except ValueError:
        return 0.0
tot_s = sum(f(d, "# Samples") for d in rows) or 1
tot_i = sum(f(d, "Instructions Executed") for d in rows) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_i))
print("%-6s %-7s %-7s %-9s %-9s %-8s %-8s %-8s %-8s  %s" % ("line", "samp%", "inst%", "shWave", "shExcess", "long_sb", "short_sb", "barrier", "mio", "source"))
key = lambda d: -f(d, "# Samples")
for d in sorted(rows, key=key)[:top]:
    print("%-6s %-7.2f %-7.2f %-9d %-9d %-8d %-8d %-8d %-8d  %s" % (
        d.get("#", d.get("Line", "?")), 100 * f(d, "# Samples") / tot_s, 100 * f(d, "Instructions Executed") / tot_i,
        f(d, "L1 Wavefronts Shared"), f(d, "L1 Wavefronts Shared Excessive"),
        f(d, "stall_long_sb"), f(d, "stall_short_sb"), f(d, "stall_barrier"), f(d, "stall_mio"), d["Source"].strip()[:110]))
