"""Aggregate the per-instruction warp-stall samples of an `ncu --page source --csv` export into the
regions between block barriers / mbarrier waits, and list the hottest instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS = hdr.index('# Samples'); iSrc = hdr.index('Source'); iEx = hdr.index('Instructions Executed')
data = []
for r in rows[2:]:
    if len(r) <= max(iS, iSrc, iEx): continue
    if r[0] == 'Address' and data: break      # a second table (source view) follows the SASS view
    try: int(r[iS] or 0)
    except ValueError: continue
    data.append(r)
tot = sum(int(r[iS] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
acc = 0; segs = []
for i, r in enumerate(data):
    acc += int(r[iS] or 0)
    src = r[iSrc]
    if 'BAR.SYNC' in src or 'SYNCS.PHASECHK' in src or 'CS2R' in src:
        segs.append((i, src.strip()[:50], acc))
prev = 0
for i, src, a in segs:
    if a - prev > 0.004 * tot:
        print(f"{i:5d} {src:52s} samples before={a-prev:7d} ({100*(a-prev)/tot:5.1f}%)")
    prev = a
print('tail', acc - prev)
for r in sorted(data, key=lambda r: -int(r[iS] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(r[iS], r[iEx], r[iSrc].strip()[:100])
