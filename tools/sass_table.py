"""Per-kernel SASS opcode table of libcrct_b200.so (cuobjdump -sass): counts of the mnemonics that show which hardware path a kernel
uses — UTCHMMA (tcgen05.mma), UTMALDG (TMA loads), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync),
LDSM (ldmatrix), LDGSTS (cp.async), MUFU.  Runs without a GPU.

    python tools/sass_table.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cqa_crct_b200', 'libcrct_b200.so')
OPS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'LDTM', 'UTCBAR', 'SYNCS', 'HMMA', 'LDSM', 'LDGSTS', 'MUFU', 'RED', 'ATOM']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.split('\n'):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            cur['_total'] += 1
            for o in OPS:
                if op == o or op.startswith(o + '.') or (o in ('RED', 'ATOM') and op.startswith(o)):
                    if o == 'UTCHMMA' and '.2CTA' in op:
                        cur['UTCHMMA.2CTA'] += 1
                    else:
                        cur[o] += 1
    rows = []
    for name, c in kernels.items():
        d = demangle(name)
        d = re.sub(r'\(anonymous namespace\)::|<unnamed>::', '', d)
        d = re.sub(r'^void ', '', d)
        d = re.sub(r'\(.*', '', d)
        rows.append((d, c))
    groups = collections.OrderedDict()               # family = kernel name without template arguments
    for d, c in rows:
        fam = re.sub(r'<.*', '', d)
        g = groups.setdefault(fam, [0, collections.Counter()])
        g[0] += 1
        g[1].update(c)
    out = ['# SASS opcode counts per kernel family of cqa_crct_b200/libcrct_b200.so (sm_100a; summed over template instantiations)',
           f'# {"kernel":44s} {"inst":>4} {"SASS":>7} ' + ' '.join(f'{o:>8}' for o in OPS)]
    for fam, (n, c) in sorted(groups.items(), key=lambda kv: -kv[1][1]['UTCHMMA'] - kv[1][1]['UTCHMMA.2CTA'] - 0.001 * kv[1][1]['HMMA']):
        out.append(f'{fam[:46]:46s} {n:>4} {c["_total"]:>7} ' + ' '.join(f'{c[o]:>8}' for o in OPS))
    out.append('')
    out.append('# per instantiation: attention and GEMM kernels')
    for d, c in rows:
        if 'attn' in d or 'gemm_tcgen05' in d:
            out.append(f'{d[:70]:70s} {c["_total"]:>6} ' + ' '.join(f'{o}={c[o]}' for o in OPS if c[o]))
    text = '\n'.join(out) + '\n'
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], 'w').write(text)


if __name__ == '__main__':
    main()
