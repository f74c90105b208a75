"""Per-source-line executed warp instructions and stall samples of one kernel launch of an .ncu-rep (built with -lineinfo).
   python profiles/src_hot.py REP LAUNCH_INDEX [min_pct=0.5]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda', '--launch-skip', str(idx),
                      '--launch-count', '1'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname = ''
h = None
lines = []
first_kernel_done = False
nk = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        nk += 1
        if nk > 1: break
        print(r[1][:110]); continue
    if r and r[0] == 'File Name': fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': h = r; continue
    if h and len(r) == len(h) and r[0].isdigit():
        iE, iN = h.index('Instructions Executed'), h.index('# Samples')
        lines.append((fname, int(r[0]), r[1].strip(), int(r[iE]) if r[iE].isdigit() else 0, int(r[iN]) if r[iN].isdigit() else 0))
totE = sum(l[3] for l in lines); totS = sum(l[4] for l in lines)
print('lines', len(lines), 'executed', totE, 'samples', totS)
for f, ln, src, e, s in lines:
    if e >= totE * minpct / 100 or s >= totS * minpct / 100:
        print('%5.1f%% %5.1f%%  %s:%d  %s' % (100 * e / totE, 100 * s / max(totS, 1), f, ln, src[:120]))
