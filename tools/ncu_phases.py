"""Per-phase digest of an `ncu --page source --csv` export of the sweeps: the SASS stream of one kernel is cut at
every block barrier / mbarrier wait inside the time loop; for every piece: warp-instructions issued per decoder step,
stall samples (share of the kernel) and the top stall reasons.  Usage: ncu_phases.py source.csv <kernel substring> [steps]"""
import csv, sys, collections
path, want = sys.argv[1], sys.argv[2]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 121
rows = list(csv.reader(open(path)))
# split into per-kernel tables
tables, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        tables.append(cur)
    elif cur is not None and r and r[0] == 'Address':
        if cur['hdr'] is None: cur['hdr'] = r
        else:
            cur = {'name': cur['name'] + ' (second view)', 'hdr': r, 'rows': []}; tables.append(cur)
    elif cur is not None and cur['hdr'] is not None:
        cur['rows'].append(r)
for tb in tables:
    if want not in tb['name'] or 'second view' in tb['name']: continue
    h = tb['hdr']
    iS, iSrc, iEx = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    data = [r for r in tb['rows'] if len(r) > iEx and r[iS].isdigit()]
    tot = sum(int(r[iS]) for r in data)
    ncta = 125
    print(tb['name']); print('total samples', tot, 'SASS instructions', len(data))
    allst = collections.Counter()
    for r in data:
        for i, c in stall_cols: allst[c] += int(r[i] or 0)
    print('kernel stall mix:', ', '.join(f"{c[6:]} {100*v/tot:.1f}%" for c, v in allst.most_common(8)))
    # loop body = instructions executed at least steps*ncta times (one warp per CTA per step)
    seg_s, seg_i, seg_st, seg_first = 0, 0, collections.Counter(), 0
    print(f"{'sass#':>6} {'cut at':46s} {'warp-instr/step/CTA':>20} {'samples %':>10}  top stalls")
    for k, r in enumerate(data):
        ex = int(r[iEx] or 0)
        s = int(r[iS])
        seg_s += s
        if ex >= steps * ncta * 0.9: seg_i += ex
        for i, c in stall_cols: seg_st[c] += int(r[i] or 0)
        src = r[iSrc].strip()
        if ('BAR.SYNC' in src or 'SYNCS.PHASECHK' in src) and ex >= steps * ncta * 0.9 or k == len(data) - 1:
            top = ', '.join(f"{c[6:]} {100*v/max(seg_s,1):.0f}%" for c, v in seg_st.most_common(4))
            print(f"{k:6d} {src[:46]:46s} {seg_i/(steps*ncta):20.0f} {100*seg_s/tot:10.1f}  {top}")
            seg_s, seg_i, seg_st = 0, 0, collections.Counter()
    print('hottest instructions:')
    for r in sorted(data, key=lambda r: -int(r[iS]))[:25]:
        st = sorted(((int(r[i] or 0), c[6:]) for i, c in stall_cols), reverse=True)[:2]
        print(f"{int(r[iS]):6d} ({100*int(r[iS])/tot:4.1f}%) ex={r[iEx]:>8} {r[iSrc].strip()[:70]:70s} {st}")
