"""Kernels launched after the backward sweep, in launch order, from an ncu launch list (durations in us)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
out = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0] != "ID"]
key = sys.argv[3] if len(sys.argv) > 3 else "dec_bwd_v3"
i0 = [i for i, d in enumerate(out) if key in d["Kernel Name"]][0]
for d in out[i0:i0 + int(sys.argv[2]) if len(sys.argv) > 2 else i0 + 40]:
    print(d["ID"], round(float(d["Metric Value"].replace(",", "")) / 1000, 1), d["Grid Size"], d["Kernel Name"][:80])
