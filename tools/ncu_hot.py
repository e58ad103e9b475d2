"""Top stall sites of one kernel from `ncu --page source --csv` (SASS view).

    ncu -i X.ncu-rep --page source --csv -k regex:<kernel> -c 1 > k.csv ; python tools/ncu_hot.py k.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ix['# Samples']] or 0) for r in body)
agg = {s: sum(int(r[ix[s]] or 0) for r in body) for s in stalls}
print('kernel:', rows[0][1][:80], ' samples', tot, ' instructions', len(body))
print('stall mix:', ', '.join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:N]
for i in sorted(top):
    r = body[i]
    s = int(r[ix['# Samples']] or 0)
    why = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * s / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:70]:70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
