"""Hot spots of an `ncu --page source --csv --print-source sass` dump: top SASS instructions by stall samples, with the dominant
stall reasons, and the totals per stall reason.  usage: python tools/ncu_source_hot.py dump.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[col['# Samples']] or 0)
    except ValueError:
        continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print(f'{len(data)} instructions, {tot} samples')
agg = {h: sum(int(r[col[h]] or 0) for _, r in data) for h in stalls}
print('by reason:', ', '.join(f'{h[6:]} {v / max(1, sum(agg.values())):.1%}' for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for k, (s, r) in enumerate(sorted(enumerate(data), key=lambda t: -t[1][0])[:top] if False else sorted(data, key=lambda t: -t[0])[:top]):
    why = sorted(((int(r[col[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f'{s:>6} {s / tot:6.1%} {r[col["Address"]][-5:]} {r[col["Source"]][:70]:<70} exec {r[col["Instructions Executed"]]:>8}  ' +
          ' '.join(f'{n}:{v}' for v, n in why if v))
