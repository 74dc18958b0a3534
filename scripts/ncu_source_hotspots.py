#!/usr/bin/env python
"""Top warp-stall sites of one kernel of an ncu report (SASS page), with the two preceding instructions:

  python scripts/ncu_source_hotspots.py gpurun_out/x.ncu-rep regex:move [N]

Reads `ncu -i ... --page source --csv -k <filter>` (works without a GPU)."""
import csv
import io
import subprocess
import sys

rep, filt = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 16
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", filt], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?")
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hi]
si, src = H.index("Warp Stall Sampling (All Samples)"), H.index("Source")
data = [(int(r[si] or 0), r[src].strip()) for r in rows[hi + 1:] if len(r) > si and r[si].isdigit()]
seen = {}
for i, (s, ins) in enumerate(data):          # the page lists the kernel once per captured launch: keep the first listing
    if i and ins == data[0][1] and s == data[0][0] and i > 8:
        data = data[:i]
        break
tot = sum(d[0] for d in data)
print(f"kernel: {name}\ninstructions {len(data)}, warp-stall samples {tot}")
for i in sorted(sorted(range(len(data)), key=lambda j: -data[j][0])[:top]):
    prev = " | ".join(d[1][:44] for d in data[max(0, i - 2):i])
    print(f"{i:5d} {100 * data[i][0] / tot:5.1f}%  {data[i][1][:52]:52s} <- {prev}")
