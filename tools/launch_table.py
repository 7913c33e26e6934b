"""Print the per-launch durations (us) of an `ncu --metrics gpu__time_duration.sum --csv` launch list, in launch
order, for the span between two consecutive forward sweeps (= one training step)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
out = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r is not hdr and r[0] != "ID"]
idx = [i for i, d in enumerate(out) if "dec_fwd_v3" in d["Kernel Name"]]
lo, hi = (idx[0], idx[1]) if len(idx) > 1 else (0, len(out))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
tot = 0.0
for d in out[lo - 30 if lo >= 30 else 0:hi - 30 if len(idx) > 1 else hi]:
    us = float(d["Metric Value"].replace(",", "")) / 1000.0
    tot += us
    if pat in d["Kernel Name"]:
        print(f"{d['ID']:>5} {us:8.1f}  {d['Grid Size']:>14}  {d['Kernel Name'][:90]}")
print(f"sum over the step: {tot:.1f} us")
