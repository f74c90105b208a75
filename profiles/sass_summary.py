"""Per-kernel SASS evidence of libbfe.so (cuobjdump -sass): counts of the mnemonics that matter for this path.
   python profiles/sass_summary.py [out.txt]      (no GPU needed)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'exptool_b200', 'libbfe.so')
KEYS = ['DMMA', 'UBLKCP', 'SYNCS', 'LDG.E.ENL2.256', 'STG.E.ENL2.256', 'LDG', 'STG', 'LDS', 'STS', 'ATOMG', 'REDG', 'ATOMS', 'MATCH',
        'DFMA', 'DMUL', 'DADD', 'MUFU', 'SHFL', 'CCTL', 'ACQBULK', 'UTMALDG', 'UTCHMMA', 'LDTM']
txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
name = None
cnt = collections.OrderedDict()
arch = set(re.findall(r'arch = (sm_\w+)', txt))
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
        cnt.setdefault(name, collections.Counter())
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        op = m.group(1)
        cnt[name]['_total'] += 1
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k.count('.') and op.startswith(k)):
                cnt[name][k] += 1
out = ['# SASS mnemonic counts per kernel of exptool_b200/libbfe.so (%s); static counts, `cuobjdump -sass`' % ', '.join(sorted(arch)),
       '# FP64 tensor path = DMMA (mma.sync.m8n8k4.f64; tcgen05 has no FP64 kind), TMA bulk copy = UBLKCP, mbarrier = SYNCS,',
       '# 256-bit global accesses = LDG/STG.E.ENL2.256; no UTCHMMA / LDTM / UTMALDG on this FP64 path (see DESIGN.md section 3.5)', '']
tot = collections.Counter()
for k, c in cnt.items():
    sel = ' '.join('%s=%d' % (q, c[q]) for q in KEYS if c[q])
    out.append('%-70s insts=%-6d %s' % (k[:70], c['_total'], sel))
    tot.update(c)
out.append('')
out.append('TOTAL kernels=%d ' % len(cnt) + ' '.join('%s=%d' % (q, tot[q]) for q in KEYS if tot[q]))
s = '\n'.join(out) + '\n'
if len(sys.argv) > 1:
    open(sys.argv[1], 'w').write(s)
print(s)
