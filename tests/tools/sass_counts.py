"""Instruction histogram of the built library per kernel (cuobjdump -sass): the tcgen05 / TMEM / TMA mnemonics that show the
sm_100a data path, plus the legacy tensor-core and SIMT ones.   python tests/tools/sass_counts.py > profiles/sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'hamer_yolo_b200', 'libhamer_b200.so')
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UBLKCP', 'UTMAPF', 'UTMACCTL', 'SYNCS', 'HMMA',
        'FFMA2', 'FFMA', 'MUFU', 'LDS', 'STS', 'LDG', 'STG', 'USETMAXREG', 'UCGABAR']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = per.setdefault(re.sub(r'\(.*', '', name).replace('void ', '').replace('hm::', ''), collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur is not None:
        op = m.group(1)
        cur['total'] += 1
        base = op.split('.')[0]
        if op.startswith('UTCHMMA') and '.2CTA' in op:
            cur['UTCHMMA.2CTA'] += 1
        if base in KEYS:
            cur[base] += 1
print(f'# {os.path.relpath(lib, ROOT)}: SASS instruction counts per kernel (static; UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,')
print('# UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store / reduce, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier)')
cols = ['total'] + KEYS
print('kernel,' + ','.join(cols))
tot = collections.Counter()
for k, c in per.items():
    print(f'"{k}",' + ','.join(str(c.get(x, 0)) for x in cols))
    tot.update(c)
print('"ALL",' + ','.join(str(tot.get(x, 0)) for x in cols))
