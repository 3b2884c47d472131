"""profiles/r2_sass_evidence.md: per-kernel counts of the Blackwell SASS mnemonics in the built library.

  python tools/sass_evidence.py > profiles/r2_sass_evidence.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mvpnet_b200', 'libmvpnet_b200.so')
PATS = collections.OrderedDict([
    ('UTCHMMA (tcgen05.mma, cta_group::1)', r'UTCHMMA(?!\.2CTA)'), ('UTCHMMA.2CTA (tcgen05.mma.cta_group::2)', r'UTCHMMA\.2CTA'),
    ('UTMALDG (TMA tensor load)', r'UTMALDG'), ('UBLKCP (bulk copy)', r'UBLKCP'), ('LDTM (tcgen05.ld)', r'\bLDTM'), ('STTM (tcgen05.st)', r'\bSTTM'),
    ('UTCBAR (tcgen05.commit)', r'UTCBAR'), ('LDGSTS (cp.async)', r'LDGSTS'), ('STG.E.ENL2.256 (256-bit store)', r'STG\.E\.ENL2\.256'),
    ('REDUX (redux.sync)', r'REDUX'), ('UCGABAR (cluster barrier)', r'UCGABAR'), ('legacy HMMA', r'\bHMMA')])


def main():
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*Function : ', txt)[1:]
    tot, rows = collections.Counter(), []
    for f in funcs:
        name = f.split('\n', 1)[0].strip()
        c = {k: len(re.findall(p, f)) for k, p in PATS.items()}
        tot.update(c)
        if any(c[k] for k in list(PATS)[:9]):
            dem = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
            rows.append((re.sub(r'\(.*', '', dem)[:70], c))
    print('# SASS evidence, round 2 (`cuobjdump -sass mvpnet_b200/libmvpnet_b200.so`, sm_100a, built by `python -m mvpnet_b200.build`; `tools/sass_evidence.py`)\n')
    print('Totals over the library (%d kernels):\n' % len(funcs))
    for k in PATS:
        print('* %s: %d' % (k, tot[k]))
    print('\n| kernel | ' + ' | '.join(k.split(' ')[0] for k in PATS) + ' |')
    print('|---|' + '---|' * len(PATS))
    for name, c in sorted(rows):
        print('| `%s` | ' % name + ' | '.join(str(c[k]) for k in PATS) + ' |')


if __name__ == '__main__':
    main()
