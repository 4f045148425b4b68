"""SASS op counts per kernel of libt3d_b200.so (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM / TMA use.
usage: python tools/sass_op_counts.py [lib.so] > profiles/rNN_sass_op_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'transferable3d_b200', 'libt3d_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'SYNCS', 'UTCCP', 'R2UR', 'F2FP', 'HMMA', 'FFMA', 'FFMA2', 'FADD2',
       'LDG', 'STG', 'ATOMG', 'REDG', 'RED']
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
name, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[name] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?\S+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1).split('.')[0]
        counts[name][op] += 1
        total[name] += 1
        if op == 'UTCHMMA' and '2CTA' in m.group(1):
            counts[name]['2CTA'] += 1
print('SASS op counts per kernel of %s (cuobjdump -sass, sm_100a); 2CTA = UTCHMMA.2CTA (cta_group::2) among the UTCHMMA' % os.path.basename(lib))
print('kernel | total | ' + ' | '.join(OPS) + ' | 2CTA')
for k in sorted(counts, key=lambda k: -counts[k]['UTCHMMA']):
    if counts[k]['UTCHMMA'] == 0 and counts[k]['UBLKCP'] == 0:
        continue
    print(k + ' | %d | ' % total[k] + ' | '.join(str(counts[k][o]) for o in OPS) + ' | %d' % counts[k]['2CTA'])
