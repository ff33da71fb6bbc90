"""Summarise an `ncu --page source --csv` dump: top stalled SASS lines with their dominant stall reason.
usage: ncu -i rep --page source --csv --kernel-id ::regex:NAME:IDX > src.csv ; python scripts/ncu_top_stalls.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(num(r[i_s]) for r in data)
print('kernel', rows[0][1][:100])
print('total samples', tot, 'warp instructions', sum(num(r[i_ex]) for r in data))
stall_cols = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
agg = {hdr[k]: sum(num(r[k]) for r in data) for k in stall_cols}
print('stall totals:', sorted(agg.items(), key=lambda t: -t[1])[:8])
for r in sorted(data, key=lambda r: -num(r[i_s]))[:n]:
    reasons = sorted(((hdr[k], num(r[k])) for k in stall_cols if num(r[k]) > 0), key=lambda t: -t[1])
    print(f'{num(r[i_s]):6d} {100*num(r[i_s])/tot:5.1f}% ex={num(r[i_ex]):8d}  {r[i_src].strip()[:64]:64s} {reasons[:2]}')
