"""Per-SASS-instruction view of one kernel of an .ncu-rep: executed count, active threads, stall samples, shared-memory
wavefronts against ideal. Usage: python tools/ncu_hotloop.py report.ncu-rep <launch index> [min share, default 0.0008]"""
import csv, subprocess, sys
rep, launch = sys.argv[1], int(sys.argv[2])
share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0008
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
isrc, iex, ith, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = [r for r in rows if len(r) > iwi and r[iex].isdigit()]
half = len(data) // 2
if half and data[0][isrc] == data[half][isrc]:
    data = data[:half]  # the page lists the function twice
tot = sum(int(r[iex]) for r in data)
print(rows[0][1][:100] if len(rows[0]) > 1 else "")
print("total warp instructions", tot, "SASS lines", len(data))
for n, r in enumerate(data):
    ex = int(r[iex])
    if ex > tot * share:
        print(f"{n:5d} {r[isrc].strip()[:62]:62s} {ex / 1e6:8.1f}M thr {int(r[ith]) // max(1, ex):2d} smp {r[ismp]:>6s} shw {int(r[iw]) / 1e6:7.1f} ideal {int(r[iwi]) / 1e6:7.1f}")
