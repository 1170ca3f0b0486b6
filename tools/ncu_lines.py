"""Per-CUDA-source-line instruction counts and stall samples of the first kernel in an .ncu-rep
(needs -lineinfo at compile time and --import-source on at capture time).  usage: ncu_lines.py rep [min_share]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iex, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows:  # (source text with quotes / commas shifts the columns: index the metrics from the right)
    if len(r) >= len(hdr) and r[0].isdigit():
        try:
            lines.append((r[0], r[1], int(r[iex - len(hdr)] or 0), int(r[ist - len(hdr)] or 0)))
        except ValueError:
            pass
tot_e, tot_s = sum(l[2] for l in lines), sum(l[3] for l in lines)
print(f"{rep}: {tot_e} warp instructions, {tot_s} stall samples")
for ln, src, e, s in lines:
    if e >= min_share * tot_e or s >= min_share * tot_s:
        print(f"{ln:>5s} {100 * e / tot_e:5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% stall  {src.strip()[:110]}")
