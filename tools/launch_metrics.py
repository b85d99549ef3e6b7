"""Per-kernel summary of one train step from an ncu launch list with several metrics:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\\
sm__warps_active.avg.pct_of_peak_sustained_active,sm__issue_active.avg.pct_of_peak_sustained_elapsed --clock-control none -s <skip> -c <n> --csv \\
        --log-file launches.csv python tools/step_profile.py 3
    python tools/launch_metrics.py launches.csv [summary.txt] [gemm_traffic.json]

One step = the launches between two consecutive adamw_kernel launches.  Durations are cold-cache and serialised (ncu replays
every kernel alone): they give each kernel's SHARE of the step and its achieved DRAM bandwidth, not the step time."""
import collections
import csv
import json
import re
import sys

HBM_PEAK = 6457.7      # GB/s, MEASURED_PEAKS.json


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    launches = collections.OrderedDict()
    for r in rows:
        k = int(r['ID'])
        d = launches.setdefault(k, {'name': r['Kernel Name']})
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        n = r['Metric Name']
        if n == 'gpu__time_duration.sum':
            v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)        # -> us
        if n.startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d[n] = v
    return list(launches.values())


def main():
    L = load(sys.argv[1])
    idx = [i for i, d in enumerate(L) if 'adamw' in d['name']]
    step = L[idx[0] + 1:idx[1] + 1] if len(idx) >= 2 else L
    agg = collections.OrderedDict()
    for d in step:
        n = re.sub(r'^void ', '', d['name'])
        n = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', n)
        n = re.sub(r'\(.*', '', n)
        a = agg.setdefault(n, collections.defaultdict(float))
        a['n'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['bytes'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
        for m, key in (('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps'),
                       ('sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue')):
            a[key] += d.get(m, 0.0) * d.get('gpu__time_duration.sum', 0.0)            # time-weighted
    tot = sum(a['us'] for a in agg.values())
    out = [f'# one train step, B = 80, packed rows: {len(step)} launches, {tot / 1e3:.2f} ms summed kernel time (ncu: cold cache, serialised)',
           f'# {"us":>9} {"share":>6} {"n":>4} {"avg us":>8} {"DRAM GB/s":>9} {"of HBM":>6} {"MB/launch":>9} {"tensor%":>7} {"warps%":>6} {"issue%":>6}  kernel']
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        gbs = a['bytes'] / (a['us'] * 1e-6) / 1e9 if a['us'] else 0.0
        out.append(f'{a["us"]:>11.1f} {a["us"] / tot:>6.1%} {int(a["n"]):>4} {a["us"] / a["n"]:>8.1f} {gbs:>9.0f} {gbs / HBM_PEAK:>6.2f} {a["bytes"] / a["n"] / 1e6:>9.2f} '
                   f'{a["tensor"] / a["us"]:>7.1f} {a["warps"] / a["us"]:>6.1f} {a["issue"] / a["us"]:>6.1f}  {n[:110]}')
    text = '\n'.join(out) + '\n'
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text)
    if len(sys.argv) > 3:
        g = [d for d in step if 'gemm_tcgen05' in d['name'] or 'gemm_wgrad_grouped' in d['name']]
        json.dump({'workload': 'train', 'gemm_launches_per_step': len(g),
                   'dram_bytes_per_launch': sum(d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0) for d in g) / max(1, len(g)),
                   'dram_bytes_read_per_step': sum(d.get('dram__bytes_read.sum', 0) for d in g), 'dram_bytes_written_per_step': sum(d.get('dram__bytes_write.sum', 0) for d in g),
                   'gemm_us_per_step_ncu': sum(d.get('gpu__time_duration.sum', 0) for d in g),
                   'how': 'ncu launch list of one eager train step (tools/step_profile.py), dram__bytes_read.sum + dram__bytes_write.sum per gemm_tcgen05 / gemm_wgrad_grouped launch'}, open(sys.argv[3], 'w'), indent=1)


if __name__ == '__main__':
    main()
