"""Per-shape table of the GEMM launches of one step (from `CRCT_BENCH_DUMP_GEMMS=prefix python bench.py`): count, mean duration,
TFLOP/s on the executed (valid-row) problem, share of the GEMM time."""
import collections
import json
import sys

EPI = {0: 'bias', 1: 'gelu', 2: 'res', 3: 'mul', 4: 'f32/wgrad', 5: 'res_f32'}
d = json.load(open(sys.argv[1]))
groups = collections.OrderedDict()
for r in d['launches']:
    M, K = r['M'], r['K']
    if r['rows'] is not None:
        if r['a_major']:
            K = min(K, r['rows'])
        else:
            M = min(M, r['rows'])
    key = (r['M'], r['N'], r['K'], r['a_major'], EPI[r['epi']])
    g = groups.setdefault(key, {'n': 0, 'us': 0.0, 'flop': 0.0})
    g['n'] += 1
    g['us'] += r['us']
    g['flop'] += 2.0 * M * r['N'] * K
tot = sum(g['us'] for g in groups.values())
print(f"workload {d['workload']}: {len(d['launches'])} GEMM launches, {tot / 1e3:.2f} ms, event-pair overhead {d['overhead_us']:.1f} us removed")
print(f"{'M(alloc)':>9} {'N':>5} {'K(alloc)':>9} {'form':>6} {'epilogue':>10} {'n':>4} {'us/launch':>10} {'TFLOP/s':>8} {'share':>6}")
for (M, N, K, am, epi), g in sorted(groups.items(), key=lambda kv: -kv[1]['us']):
    print(f"{M:>9} {N:>5} {K:>9} {'wgrad' if am else 'fwd/dg':>6} {epi:>10} {g['n']:>4} {g['us'] / g['n']:>10.1f} {g['flop'] / g['us'] / 1e6:>8.0f} {g['us'] / tot:>6.1%}")
