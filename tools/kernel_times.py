"""Mean duration per kernel name from an `ncu --metrics gpu__time_duration.sum --csv` log (stdin or file)."""
import collections, csv, re, sys
lines = [l for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin) if l.startswith('"')]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    n = re.sub(r'\(.*', '', re.sub(r'^void |<unnamed>::|\(anonymous namespace\)::', '', r['Kernel Name']))
    v = float(r['Metric Value'].replace(',', ''))
    v = v / 1e3 if r['Metric Unit'] in ('ns', 'nsecond') else v
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += v
for n, (c, t) in agg.items():
    print(f'{n[:60]:60s} n={c:4d} mean {t / c:8.2f} us  total {t / 1e3:7.3f} ms')
