"""Per-instruction view of one kernel launch of an .ncu-rep (source page, SASS): executed warp instructions and stall samples,
grouped in blocks of consecutive instructions so the hot regions show up.
   python profiles/sass_hot.py REP LAUNCH_INDEX [block=40]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', str(idx),
                      '--launch-count', '1'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:100])
h = rows[1]
iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
body = [r for r in rows[2:] if len(r) > iN and r[iE].isdigit()]
totE = sum(int(r[iE]) for r in body); totS = sum(int(r[iN]) for r in body)
print('instructions', len(body), 'executed', totE, 'samples', totS)
for b0 in range(0, len(body), blk):
    part = body[b0:b0 + blk]
    e = sum(int(r[iE]) for r in part); s = sum(int(r[iN]) for r in part)
    ops = {}
    for r in part:
        op = r[iS].split()[0] if not r[iS].strip().startswith('@') else r[iS].split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + int(r[iE])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print('%5d-%5d  exec %5.1f%%  samples %5.1f%%  %s' % (b0, b0 + len(part), 100 * e / totE, 100 * s / max(totS, 1),
                                                         ' '.join('%s:%d' % (k, v * 100 // max(e, 1)) for k, v in top)))
