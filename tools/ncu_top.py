"""Top stalled SASS instructions of an .ncu-rep (source page):  python tools/ncu_top.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
isrc, iss, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iss] or 0) for r in data)
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0) + int(r[i] or 0)
print("samples", tot, "instructions", len(data), "warp-instr executed", sum(int(r[iex] or 0) for r in data))
print("by reason:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:top_n]:
    st = sorted(((hdr[i][6:], int(r[i] or 0)) for i in stall_cols if int(r[i] or 0) > 0), key=lambda kv: -kv[1])[:2]
    print(r[iss].rjust(6), r[iex].rjust(8), r[isrc][:72].ljust(72), st)
