"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for one train step
(the launches between two consecutive adamw_kernel launches)."""
import collections
import csv
import re
import sys


def summarize(path, out=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [r['Kernel Name'] for r in rows]
    idx = [i for i, n in enumerate(names) if 'adamw' in n]
    step = rows[idx[0] + 1:idx[1] + 1] if len(idx) >= 2 else rows
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in step:
        n = re.sub(r'\(.*', '', r['Kernel Name'])
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v for c, v in agg.values())
    text = [f'# one train step B=80: {len(step)} launches, {tot / 1e3:.2f} ms summed kernel time (ncu, cold-cache, serialised)']
    for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        text.append(f'{v:10.1f} us {100 * v / tot:5.1f}% n={c:4d} avg={v / c:8.1f} us  {n[:120]}')
    s = '\n'.join(text) + '\n'
    if out:
        open(out, 'w').write(s)
    return s


if __name__ == '__main__':
    print(summarize(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None))
