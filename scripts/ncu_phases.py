"""Aggregate an `ncu --page source --csv` dump per kernel phase (split at barriers / mbarrier waits).
usage: python scripts/ncu_phases.py file.csv [kernel-index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
sections = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        sections.append({'name': r[1], 'hdr': None, 'data': []})
    elif sections and sections[-1]['hdr'] is None and 'Source' in r:
        sections[-1]['hdr'] = r
    elif sections and sections[-1]['hdr'] is not None and len(r) == len(sections[-1]['hdr']):
        sections[-1]['data'].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else len(sections) - 1
sec = sections[which]; hdr = sec['hdr']; data = sec['data']
print(len(sections), 'kernels; showing', which, sec['name'][:80])
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data); totS = sum(int(r[iN]) for r in data)
print('total warp-inst', tot, 'samples', totS, 'sass lines', len(data))
acc = accS = start = 0; mix = {}
def flush(i, why):
    global acc, accS, start, mix
    top = sorted(mix.items(), key=lambda x: -x[1])[:9]
    print(f'[{start:4d},{i:4d}) inst {acc/tot*100:5.1f}% smp {accS/totS*100:5.1f}% end:{why[:26]:26s}', ' '.join(f'{k}:{v/tot*100:.1f}' for k, v in top))
    acc = accS = 0; start = i; mix = {}
for i, r in enumerate(data):
    s = r[iS].strip(); t = s.split()
    op = (t[1] if s.startswith('@') else t[0]).split('.')[0]
    e = int(r[iE]); acc += e; accS += int(r[iN]); mix[op] = mix.get(op, 0) + e
    if 'BAR.SYNC' in s or 'SYNCS.PHASECHK' in s or s.startswith('EXIT') or s.startswith('RET'):
        flush(i + 1, s)
flush(len(data), 'end')
