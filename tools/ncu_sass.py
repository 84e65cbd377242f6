"""SASS-level hot spots from an ncu report.  usage: python tools/ncu_sass.py report.ncu-rep kernel_regex [topN]"""
import csv, io, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows, hdr, nk = [], None, 0
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        nk += 1
        continue
    if nk > 1:
        break
    if row[0] == "Address":
        hdr = row; continue
    if hdr and len(row) == len(hdr):
        rows.append(dict(zip(hdr, row)))
def f(d, k):
    try: return float(d.get(k, "0").replace(",", "") or 0)
    except ValueError: return 0.0
S = sum(f(d, "# Samples") for d in rows) or 1
I = sum(f(d, "Instructions Executed") for d in rows) or 1
print("instructions (static) %d, warp-instr executed %d, samples %d" % (len(rows), I, S))
ops = collections.defaultdict(lambda: [0, 0, 0, 0])
for d in rows:
    op = d["Source"].split()[0] if not d["Source"].startswith("@") else d["Source"].split()[1]
    op = op.split(".")[0] + ("." + ".".join(x for x in d["Source"].split()[0 if not d["Source"].startswith("@") else 1].split(".")[1:] if x in ("64","128","32","E","SYNC")) if op in ("LDS","STS","LDG","STG") else "")
    o = ops[op]; o[0] += f(d, "Instructions Executed"); o[1] += f(d, "# Samples"); o[2] += f(d, "L1 Wavefronts Shared"); o[3] += f(d, "L1 Wavefronts Shared Excessive")
print("%-16s %8s %8s %12s %12s" % ("opcode", "inst%", "samp%", "shWave", "shExcess"))
for op, o in sorted(ops.items(), key=lambda kv: -kv[1][0])[:22]:
    print("%-16s %8.2f %8.2f %12d %12d" % (op, 100 * o[0] / I, 100 * o[1] / S, o[2], o[3]))
print()
print("%-5s %-7s %-7s %-9s %-9s %-7s %-7s %-7s %-7s %-7s  %s" % ("idx", "samp%", "inst%", "shWave", "shExcess", "long", "short", "barr", "mio", "wait", "sass"))
for i, d in sorted(enumerate(rows), key=lambda kv: -f(kv[1], "# Samples"))[:top]:
    print("%-5d %-7.2f %-7.2f %-9d %-9d %-7d %-7d %-7d %-7d %-7d  %s" % (
        i, 100 * f(d, "# Samples") / S, 100 * f(d, "Instructions Executed") / I, f(d, "L1 Wavefronts Shared"),
        f(d, "L1 Wavefronts Shared Excessive"), f(d, "stall_long_sb"), f(d, "stall_short_sb"), f(d, "stall_barrier"),
        f(d, "stall_mio"), f(d, "stall_wait"), d["Source"][:90]))
if len(sys.argv) > 4:
    lo, hi = map(int, sys.argv[4].split(":"))
    for i in range(lo, hi):
        d = rows[i]
        print("%-5d %-6.2f sh=%d/%d  %s" % (i, 100 * f(d, "# Samples") / S, f(d, "L1 Wavefronts Shared"), f(d, "L1 Wavefronts Shared Excessive"), d["Source"][:100]))
