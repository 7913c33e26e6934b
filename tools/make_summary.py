"""Regenerate the per-kernel table of profiles/r01_v5_summary.md from an ncu launch list + bench line + chain log.
usage: python tools/make_summary.py profiles/r01_v5_launches.csv profiles/r01_v5_bench.json chain.log"""
import collections, csv, json, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(r for r in rows if "Kernel Name" in r)
out = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r[0] != "ID"]
fw = [i for i, d in enumerate(out) if "dec_fwd_v3" in d["Kernel Name"]]
step = out[fw[0]:fw[1]]          # one training step (sweep to sweep)
agg = collections.OrderedDict()
for d in step:
    n = d["Kernel Name"]
    n = (n[5:] if n.startswith("void ") else n).split("(")[0]
    if "group_tn" in n and d["Grid Size"].startswith("(23,"):
        n += " [shadow launches, 23 CTAs]"
    a = agg.setdefault(n, [0.0, 0])
    a[0] += float(d["Metric Value"].replace(",", "")) / 1000
    a[1] += 1
tot = sum(a[0] for a in agg.values())
b = json.load(open(sys.argv[2]))
print("| total us | launches | share | kernel |\n|---|---|---|---|")
for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {us:.1f} | {c} | {100 * us / tot:.1f}% | `{n[:90]}` |")
print(f"\nSum over the step (serialised): {tot:.0f} us in {len(step)} launches.  Bench line of the same build: "
      f"{b['value']:.0f} ex/s, {b['ms_per_step']:.3f} ms/step, e2e {b['e2e']['value']:.0f} ex/s, "
      f"{b['gpu_launches']} launches of our kernels in the {b['steps']} timed steps.")
if len(sys.argv) > 3:
    print("\n```\n" + "\n".join(open(sys.argv[3]).read().strip().splitlines()[-2:]) + "\n```")
